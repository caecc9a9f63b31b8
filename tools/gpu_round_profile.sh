#!/bin/bash
# Round artefacts: default bench line, ncu launch list of the same command, ncu --set full of the top kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${R:-r1}
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
# launch list: the timed region of the same command (graph replay), one step
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sample --profiler-range > gpurun_out/${R}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${R}_launches.csv 30 | tee gpurun_out/${R}_launch_summary.txt
# full captures (eager step so that launch indices are stable): level-1 instances of the heaviest kernels
ARGS="bench.py --profile-step --no-graphs --warmup 3"
COMMON="--profile-from-start off --set full --import-source on --clock-control none -f"
timeout 600 ncu $COMMON -k regex:gemm_tc_kernel -s 485 -c 1 -o gpurun_out/${R}_dgrad2 python $ARGS > gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:wgrad_tc_kernel -s 195 -c 1 -o gpurun_out/${R}_wgrad2 python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/${R}_conv2 python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_coupling_kernel|rows_mix_kernel|im2col_rows_warp" -c 3 -o gpurun_out/${R}_flow_fwd python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd|rows_coupling_bwd" -s 140 -c 2 -o gpurun_out/${R}_flow_bwd python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
grep -E "Report|rror" gpurun_out/${R}_ncu_full.log
ls -la gpurun_out/${R}_*.ncu-rep
