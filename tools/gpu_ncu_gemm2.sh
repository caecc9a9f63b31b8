#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-n2}
for kind in ${KINDS:-bwd3 bwd fwd}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_$kind python tools/bench_gemm.py $kind 1x1 262144 > gpurun_out/${TAG}_ncu_$kind.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_$kind.log
done
ls -la gpurun_out/${TAG}_*.ncu-rep
