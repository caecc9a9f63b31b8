#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -f -o gpurun_out/g2_fwd python tools/bench_gemm.py fwd 1x1 65536 > gpurun_out/g2_ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 3 -c 2 -f -o gpurun_out/g2_wgrad python tools/bench_gemm.py wgrad 1x1 65536 > gpurun_out/g2_ncu_wgrad.log 2>&1
tail -3 gpurun_out/g2_ncu_fwd.log gpurun_out/g2_ncu_wgrad.log
ls -la gpurun_out/*.ncu-rep
