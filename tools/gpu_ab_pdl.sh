#!/bin/bash
# A/B of programmatic dependent launch inside the captured training step: off / early trigger / late GEMM trigger
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-p4}
B=${B:-512}
: > gpurun_out/${TAG}_pdl.log
run() { echo "== $1 $2" >> gpurun_out/${TAG}_pdl.log; env $1 $2 timeout 400 python bench.py --batch $B --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn >> gpurun_out/${TAG}_pdl.log; }
run GLOWK_PDL=0 X=0
run GLOWK_PDL=1 X=0
run GLOWK_PDL=1 GLOWK_LIB=$PWD/pytorch_glow_b200/libglowk_pdl_late.so
run GLOWK_PDL=0 X=1
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_pdl.log" % os.environ.get("TAG","p4")):
    if l.startswith("{"):
        d=json.loads(l); print("   ", d["config"]["per_gpu_batch"], "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"]), "sm_mhz", d["clocks"]["sm_mhz"])
    else: print(l.strip()[:300])
PY
