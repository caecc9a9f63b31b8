#!/bin/bash
# full GPU test suite, then A/B of an environment switch on the short bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-ab}
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -n 8 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_bench.log
for cfg in ${CONFIGS:-"GLOWK_PDL=1:128" "GLOWK_PDL=0:128" "GLOWK_PDL=1:256"}; do
  envs=${cfg%%:*}; b=${cfg##*:}
  echo "== $envs batch $b" >> gpurun_out/${TAG}_bench.log
  env $envs timeout 400 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn >> gpurun_out/${TAG}_bench.log
done
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_bench.log" % os.environ.get("TAG","ab")):
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["per_gpu_batch"], "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"]), "roof", round(d["roofline"]["frac"],3), "loss", d["loss_bits_per_dim"])
    else: print(l.strip()[:300])
PY
