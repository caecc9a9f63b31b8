#!/bin/bash
# Round-2 artefacts of the final code: full GPU test suite, smoke(), default bench line + reference arm, and
# ncu --set full of the level-1 fused kernels of one eager training step (B = 512).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${R:-r2z}
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${R}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${R}_smoke.txt
timeout 900 python bench.py > gpurun_out/${R}_bench_B512.json 2> gpurun_out/${R}_bench.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/${R}_bench_B512.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
cut -c1-400 gpurun_out/${R}_bench_reference.json
ARGS="bench.py --profile-step --no-graphs --warmup 3"
COMMON="--profile-from-start off --set full --import-source on --clock-control none -f"
# launch order of cnet_chain_kernel in one eager step: 32 x level-1 forward, 32 x level-2 forward, 32 x level-2 backward,
# 32 x level-1 backward
timeout 600 ncu $COMMON -k regex:cnet_chain_kernel -s 0 -c 2 -o gpurun_out/${R}_cnet_fwd_train python $ARGS > gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:cnet_chain_kernel -s 96 -c 2 -o gpurun_out/${R}_cnet_bwd python $ARGS >> gpurun_out/${R}_ncu_full.log 2>&1
grep -E "Report|rror" gpurun_out/${R}_ncu_full.log
ls -la gpurun_out/${R}_*.ncu-rep
