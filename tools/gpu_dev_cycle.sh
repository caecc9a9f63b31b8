#!/bin/bash
# development cycle: full GPU tests, eager launch list at B=512, graph bench, ncu --set full of mix_bwd / coupling at levels 3 and 2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-dev}
B=${B:-512}
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -n 6 gpurun_out/${TAG}_tests.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-step --no-graphs --warmup 3 --batch $B > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv 30
timeout 400 python bench.py --batch $B --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn > gpurun_out/${TAG}_bench.log
python - <<'PY'
import json,os
for l in open("gpurun_out/%s_bench.log" % os.environ.get("TAG","dev")):
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["per_gpu_batch"], "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"]), "roof", round(d["roofline"]["frac"],3), "loss", d["loss_bits_per_dim"])
        for r in d["roofline_all"]: print("   ", r["id"], round(r["us_per_launch"],1), "us frac", round(r["frac"],3))
    else: print(l.strip()[:300])
PY
if [ -z "$NONCU" ]; then
ARGS="bench.py --profile-step --no-graphs --warmup 3 --batch $B"
COMMON="--profile-from-start off --set full --import-source on --clock-control none -f"
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd" -s 0 -c 1 -o gpurun_out/${TAG}_mixbwd_l3 python $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd" -s 40 -c 1 -o gpurun_out/${TAG}_mixbwd_l2 python $ARGS >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_coupling_kernel" -s 40 -c 1 -o gpurun_out/${TAG}_coupling_l2 python $ARGS >> gpurun_out/${TAG}_ncu_full.log 2>&1
grep -E "Report|rror" gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_*.ncu-rep
fi
