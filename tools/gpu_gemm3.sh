#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-g3}
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3 $4" >> gpurun_out/${TAG}_sweep.log; env $1 timeout 120 python tools/bench_gemm.py $2 $3 $4 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for m in 262144 65536; do
  run GLOWK_GEMM_DEBUG=0 bwd 1x1 $m
  run GLOWK_GEMM_DEBUG=32 bwd 1x1 $m
  run GLOWK_GEMM_DEBUG=0 bwd3 1x1 $m
  run GLOWK_GEMM_DEBUG=32 bwd3 1x1 $m
done
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 40
