#!/bin/bash
# cycle h2: full GPU tests, RELU_BWD micro-benchmark with / without the g*y column sum, graph bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-h2}
B=${B:-512}
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -n 15 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_sweep.log
run() { echo "== $1 $2 $3" >> gpurun_out/${TAG}_sweep.log; env $1 timeout 60 python tools/bench_gemm.py $2 0x0 $3 >> gpurun_out/${TAG}_sweep.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/${TAG}_sweep.log; }
for kind in bwd bwd3; do
  run GLOWK_GEMM_DEBUG=0 $kind 524288
  run NOGY=1 $kind 524288
  run NOGY=1 $kind 131072
done
grep -v Warning gpurun_out/${TAG}_sweep.log | tail -n 20
timeout 400 python bench.py --batch $B --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn > gpurun_out/${TAG}_bench.log
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_bench.log" % os.environ.get("TAG","h2")):
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["per_gpu_batch"], "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"]), "roof", round(d["roofline"]["frac"],3), "loss", d["loss_bits_per_dim"])
        for r in d["roofline_all"]: print("   ", r["id"], round(r["us_per_launch"],1), "us frac", round(r["frac"],3))
    else: print(l.strip()[:300])
PY
