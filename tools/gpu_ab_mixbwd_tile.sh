#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-h9}
for kb in 72 110 150; do
GLOWK_MIXBWD_SMEM_KB=$kb timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches_$kb.csv python bench.py --profile-step --no-graphs --warmup 3 --batch 512 > gpurun_out/${TAG}_ncu.log 2>&1
echo "== smem cap $kb KB"; python tools/summarize_launches.py gpurun_out/${TAG}_launches_$kb.csv 30 | grep -E "total|mix_bwd|coupling"
done
