#!/bin/bash
# GPU pass 2: full gpu test-suite, first bench lines, ncu launch list of one train step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-graphs --no-sample --no-cpu-baseline > gpurun_out/r2_bench_nograph.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r2_launches.csv python bench.py --profile-step --no-graphs --warmup 3 > gpurun_out/r2_ncu.log 2>&1
for f in gpurun_out/r2_*.log; do echo "=== $f"; tail -n 8 "$f"; done
