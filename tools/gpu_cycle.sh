#!/bin/bash
# one dev cycle: selected tests, launch-list profile of one eager step, a short graph-mode bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-c}
B=${B:-128}
timeout 900 python -m pytest ${TESTS:-tests/test_gpu_rows.py} -m gpu -q --tb=short -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -n 12 gpurun_out/${TAG}_tests.log
if [ -z "$NOPROF" ]; then
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-step --no-graphs --warmup 3 --batch $B > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv 24
fi
timeout 400 python bench.py --batch $B --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn > gpurun_out/${TAG}_bench.log
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_bench.log" % os.environ.get("TAG","c")):
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["per_gpu_batch"], "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"]), "roof", round(d["roofline"]["frac"],3), "loss", d["loss_bits_per_dim"])
    else: print(l.strip()[:300])
PY
