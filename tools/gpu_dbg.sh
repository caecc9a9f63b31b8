#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/g3_dbg.log
for d in 12 14 8; do
  echo "== GLOWK_GEMM_DEBUG=$d" >> gpurun_out/g3_dbg.log
  GLOWK_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py fwd 1x1 262144 >> gpurun_out/g3_dbg.log 2>&1
  GLOWK_GEMM_DEBUG=$d timeout 120 python tools/bench_gemm.py fwd 1x1 65536 >> gpurun_out/g3_dbg.log 2>&1
done
grep -v Warn gpurun_out/g3_dbg.log
