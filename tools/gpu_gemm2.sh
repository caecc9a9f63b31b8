#!/bin/bash
# GEMM experiments: kernel tests, then micro-bench variants (env / cluster) at M=262144
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-g2}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_kernels.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_kernels.log
tail -n 4 gpurun_out/${TAG}_kernels.log
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3 $4" >> gpurun_out/${TAG}_sweep.log; env $1 timeout 120 python tools/bench_gemm.py $2 $3 $4 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for m in 262144; do
  run GLOWK_GEMM_DEBUG=0 bwd 1x1 $m
  run GLOWK_GEMM_DEBUG=16 bwd 1x1 $m
  run GLOWK_GEMM_DEBUG=0 bwd3 1x1 $m
  run GLOWK_GEMM_DEBUG=16 bwd3 1x1 $m
  for cl in 2x1 4x1 8x1 2x2 4x2; do run GLOWK_GEMM_DEBUG=0 fwd $cl $m; done
  for cl in 4x1 8x1; do run GLOWK_GEMM_DEBUG=0 bwd $cl $m; done
done
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 40
