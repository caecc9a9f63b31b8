#!/bin/bash
# A/B of TMA-multicast clusters for every tcgen05 GEMM of the step (GLOWK_GEMM_CLUSTER), same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-h18}
: > gpurun_out/${TAG}_cluster.log
run() { echo "== $1" >> gpurun_out/${TAG}_cluster.log; env $1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sample 2>&1 | grep -v Warn >> gpurun_out/${TAG}_cluster.log; }
run X=0
run GLOWK_GEMM_CLUSTER=2x1
run GLOWK_GEMM_CLUSTER=1x2
run X=1
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_cluster.log" % os.environ.get("TAG","h18")):
    if l.startswith("{"):
        d=json.loads(l); print("   train", round(d["value"]), "ms", round(d["ms_per_step"],2), "sm_mhz", d["clocks"]["sm_mhz"], " ".join("%s=%.0f" % (r["id"], r["us_per_launch"]) for r in d["roofline_all"]))
    else: print(l.strip()[:200])
PY
