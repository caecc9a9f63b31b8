#!/bin/bash
# resident-B GEMM check: kernel tests (trap-safe), A/B micro-bench at M=262144, model tests, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-rb}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_kernels.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_kernels.log
tail -n 6 gpurun_out/${TAG}_kernels.log
: > gpurun_out/${TAG}_sweep.log
for r in 1 0; do
  echo "== GLOWK_GEMM_RESB=$r" >> gpurun_out/${TAG}_sweep.log
  for kind in fwd bwd c1 c3; do
    for m in 262144 16384; do
      GLOWK_GEMM_RESB=$r timeout 120 python tools/bench_gemm.py $kind 1x1 $m >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "$kind $m FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log
    done
  done
done
timeout 120 python tools/bench_gemm.py wgrad 1x1 262144 >> gpurun_out/${TAG}_sweep.log 2>&1
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 40
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_backward.py tests/test_gpu_rows.py -m gpu -q --tb=short > gpurun_out/${TAG}_model.log 2>&1
tail -n 4 gpurun_out/${TAG}_model.log
for r in 1 0; do
GLOWK_GEMM_RESB=$r timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESB=$r train', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'sample', round(d['sample']['value']), 'loss', d['loss_bits_per_dim']); [print('   ', r['id'], round(r['us_per_launch'],1), 'us frac', round(r['frac'],3)) for r in d['roofline_all']]
"
done
