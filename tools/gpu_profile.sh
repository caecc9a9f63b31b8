#!/bin/bash
# ncu launch list of ONE eager training step after warm-up (+ optional extra pytest file first).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-p}
B=${B:-128}
if [ -n "$PRETEST" ]; then timeout 600 python -m pytest $PRETEST -m gpu -q --tb=short -x > gpurun_out/${TAG}_pretest.log 2>&1; tail -n 8 gpurun_out/${TAG}_pretest.log; fi
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-step --no-graphs --warmup 3 --batch $B > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv 40
