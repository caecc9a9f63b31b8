#!/bin/bash
# rows-path bring-up: new kernel tests under compute-sanitizer-free run, then full GPU suite, then bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-w1}
timeout 900 python -m pytest tests/test_gpu_rows.py -m gpu -q --tb=short -x > gpurun_out/${TAG}_rows.log 2>&1
echo "rows rc=$?" >> gpurun_out/${TAG}_rows.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_rows.py > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
for b in ${BATCHES:-64 128}; do
  timeout 400 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn > gpurun_out/${TAG}_bench_$b.log
done
tail -n 25 gpurun_out/${TAG}_rows.log; tail -n 15 gpurun_out/${TAG}_tests.log
python - <<'PY'
import json,os,glob
for f in sorted(glob.glob("gpurun_out/%s_bench_*.log" % os.environ.get("TAG","w1"))):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(d["config"]["per_gpu_batch"], "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"]), "roof", round(d["roofline"]["frac"],3))
        else: print(l.strip()[:300])
PY
