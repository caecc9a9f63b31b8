#!/bin/bash
# cycle h5: ncu --set full of the window mix adjoint at levels 3 and 1; the other BASELINE configs through bench.py
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-h5}
ARGS="bench.py --profile-step --no-graphs --warmup 3 --batch 512"
COMMON="--profile-from-start off --set full --import-source on --clock-control none -f"
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd" -s 0 -c 1 -o gpurun_out/${TAG}_mixbwd_l3 python $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu $COMMON -k regex:"rows_mix_bwd" -s 70 -c 1 -o gpurun_out/${TAG}_mixbwd_l1 python $ARGS >> gpurun_out/${TAG}_ncu_full.log 2>&1
grep -E "Report|rror" gpurun_out/${TAG}_ncu_full.log
for wl in cifar32 cifar32_additive; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn > gpurun_out/${TAG}_bench_$wl.log
done
timeout 900 python bench.py --workload celebahq256 --steps 3 --warmup 3 --no-cpu-baseline --sample-batch 32 2>&1 | grep -v Warn > gpurun_out/${TAG}_bench_celebahq256.log
python - <<'PY'
import json,os,glob
for f in sorted(glob.glob("gpurun_out/%s_bench_*.log" % os.environ.get("TAG","h5"))):
    print(f)
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print("   ", d["config"]["workload"][:40], d["config"]["per_gpu_batch"], "train", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), "sample", d["sample"] and round(d["sample"]["value"],1), "loss", d["loss_bits_per_dim"])
        else: print("   ", l.strip()[:300])
PY
