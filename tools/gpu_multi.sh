#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
TAG=${TAG:-n2}
nvidia-smi -L > gpurun_out/${TAG}_gpus.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_bench.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_ref.log
grep -v -i warn gpurun_out/${TAG}_bench.log | tail -5 | cut -c1-1500
grep -v -i warn gpurun_out/${TAG}_ref.log | tail -3 | cut -c1-600
