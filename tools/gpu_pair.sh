#!/bin/bash
# cta_group::2 pair-mode GEMM vs single-CTA kernel with per-role wait counters (GLOWK_GEMM_DEBUG=64)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-pr}
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3 $4" >> gpurun_out/${TAG}_sweep.log; env $1 $2 timeout 60 python tools/bench_gemm.py $3 0x0 $4 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for kind in fwd bwd bwd3 c3; do
  for m in 262144; do
    run GLOWK_GEMM_PAIR=0 GLOWK_GEMM_DEBUG=64 $kind $m
    run GLOWK_GEMM_PAIR=1 GLOWK_GEMM_DEBUG=64 $kind $m
  done
done
run GLOWK_GEMM_PAIR=1 GLOWK_GEMM_DEBUG=0 fwd 640
run GLOWK_GEMM_PAIR=1 GLOWK_GEMM_DEBUG=0 bwd 16384
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 40
