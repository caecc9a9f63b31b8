#!/bin/bash
# First GPU pass: SIMT/fp32 parity first, tcgen05 in separate processes (a trap must not poison the rest).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "not tcgen05 and not relu_bwd and not wgrad" > gpurun_out/t1_kernels.log 2>&1
python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -k "not bf16 and not full_size" > gpurun_out/t1_model_fp32.log 2>&1
timeout 300 python tools/debug_tc.py gemm > gpurun_out/t1_debug_gemm.log 2>&1
timeout 300 python tools/debug_tc.py wgrad > gpurun_out/t1_debug_wgrad.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "tcgen05" > gpurun_out/t1_tc.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "relu_bwd or wgrad" > gpurun_out/t1_tc_bwd.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -s -k "bf16 or full_size" > gpurun_out/t1_model_bf16.log 2>&1
for f in gpurun_out/t1_*.log; do echo "=== $f"; tail -n 12 "$f"; done
