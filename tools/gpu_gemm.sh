#!/bin/bash
# GEMM rewrite check: kernel tests first, then the cluster sweep (one process per case, bounded).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x > gpurun_out/g1_kernels.log 2>&1
echo "rc=$?" >> gpurun_out/g1_kernels.log
tail -n 15 gpurun_out/g1_kernels.log
: > gpurun_out/g1_sweep.log
for kind in fwd bwd c1 c3; do
  for cl in ${CLUSTERS:-1x1 2x1 1x2}; do
    timeout 120 python tools/bench_gemm.py $kind $cl >> gpurun_out/g1_sweep.log 2>&1 || echo "$kind $cl FAILED rc=$?" >> gpurun_out/g1_sweep.log
  done
done
timeout 120 python tools/bench_gemm.py wgrad >> gpurun_out/g1_sweep.log 2>&1
grep -v Warning gpurun_out/g1_sweep.log | tail -n 90
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_backward.py -m gpu -q --tb=short > gpurun_out/g1_model.log 2>&1
tail -n 8 gpurun_out/g1_model.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/g1_bench.log 2>&1
tail -n 3 gpurun_out/g1_bench.log
