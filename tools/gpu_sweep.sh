#!/bin/bash
# Per-GPU batch sweep of the training bench (no CPU baseline / sampling legs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-s1}
: > gpurun_out/${TAG}_sweep.log
for b in ${BATCHES:-64 128 256}; do
  timeout 300 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-sample 2>&1 | grep -v Warn >> gpurun_out/${TAG}_sweep.log
  nvidia-smi --query-gpu=memory.used --format=csv,noheader >> gpurun_out/${TAG}_sweep.log
done
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_sweep.log" % os.environ.get("TAG","s1")):
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["per_gpu_batch"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"])
    else: print(l.strip()[:200])
PY
