#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-h7}
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -n 8 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3 $4" >> gpurun_out/${TAG}_sweep.log; env $1 $2 timeout 60 python tools/bench_gemm.py $3 0x0 $4 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for kind in bwd bwd3; do
  run X=1 Y=1 $kind 524288
  run NOGY=1 Y=1 $kind 524288
  run NOGY=1 NOGB=1 $kind 524288
  run NOGY=1 Y=1 $kind 524288
  run NOGY=1 NOGB=1 $kind 524288
done
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 24
