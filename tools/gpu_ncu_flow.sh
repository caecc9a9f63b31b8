#!/bin/bash
# ncu --set full captures of level-1 instances of the rows-path flow kernels (one eager train step, B=128)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-f}
B=${B:-128}
ARGS="bench.py --profile-step --no-graphs --warmup 3 --batch $B"
COMMON="--profile-from-start off --set full --import-source on --clock-control none -f"
timeout 600 ncu $COMMON -k regex:"rows_mix_kernel|im2col_rows_warp|rows_coupling_kernel" -c 3 -o gpurun_out/${TAG}_flow_fwd python $ARGS > gpurun_out/${TAG}_ncu_flow.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd|rows_coupling_bwd" -s 140 -c 2 -o gpurun_out/${TAG}_flow_bwd python $ARGS >> gpurun_out/${TAG}_ncu_flow.log 2>&1
grep -E "PROF|rror" gpurun_out/${TAG}_ncu_flow.log | tail -12
ls -la gpurun_out/${TAG}_flow*.ncu-rep
