#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-e4}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "relu_bwd or tcgen05 or gemm" > gpurun_out/${TAG}_kernels.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_kernels.log
tail -n 12 gpurun_out/${TAG}_kernels.log
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3" >> gpurun_out/${TAG}_sweep.log; env $1 timeout 60 python tools/bench_gemm.py $2 0x0 $3 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for kind in bwd bwd3; do
  run GLOWK_GEMM_DEBUG=64 $kind 262144
  run GLOWK_GEMM_DEBUG=0 $kind 16384
  run GLOWK_GEMM_DEBUG=0 $kind 640
done
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 30
