#!/bin/bash
# Round artefacts: default bench line + reference arm, ncu launch list of the same command (graph replay), eager
# launch list, ncu --set full of the top kernels (level-1 instances, B=512).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${R:-r1h}
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
cut -c1-400 gpurun_out/${R}_bench_reference.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${R}_launches_graph.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sample --profiler-range > gpurun_out/${R}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${R}_launches_graph.csv 32 | tee gpurun_out/${R}_launches_graph_summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${R}_launches_eager.csv python bench.py --profile-step --no-graphs --warmup 3 > gpurun_out/${R}_ncu_launches2.log 2>&1
python tools/summarize_launches.py gpurun_out/${R}_launches_eager.csv 32 > gpurun_out/${R}_launches_eager_summary.txt
ARGS="bench.py --profile-step --no-graphs --warmup 3"
COMMON="--profile-from-start off --set full --import-source on --clock-control none -f"
timeout 600 ncu $COMMON -k regex:gemm_tc_kernel -s 485 -c 1 -o gpurun_out/${R}_dgrad2 python $ARGS > gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:gemm_tc_kernel -s 484 -c 1 -o gpurun_out/${R}_dgrad3 python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:wgrad_tc_kernel -s 195 -c 1 -o gpurun_out/${R}_wgrad2 python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/${R}_conv2 python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_coupling_kernel|rows_mix_kernel|im2col_rows_piece" -c 3 -o gpurun_out/${R}_flow_fwd python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd|rows_coupling_bwd" -s 140 -c 2 -o gpurun_out/${R}_flow_bwd python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
grep -E "Report|rror" gpurun_out/${R}_ncu_full.log
ls -la gpurun_out/${R}_*.ncu-rep
