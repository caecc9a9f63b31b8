#!/bin/bash
# ncu launch list of ONE eager training step (and optionally one sampling pass) after warm-up.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-p}
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-step --no-graphs --warmup 3 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv 40
