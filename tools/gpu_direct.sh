#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-dr}
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3" >> gpurun_out/${TAG}_sweep.log; env $1 timeout 60 python tools/bench_gemm.py $2 0x0 $3 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for kind in fwd c1 bwd bwd3; do
  run GLOWK_GEMM_DEBUG=64 $kind 262144
  run GLOWK_GEMM_DEBUG=192 $kind 262144
done
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 40
