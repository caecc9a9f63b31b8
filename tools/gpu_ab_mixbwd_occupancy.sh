#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-h21}
prof() { env $2 $3 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches_$1.csv python bench.py --profile-step --no-graphs --warmup 3 --batch 512 > gpurun_out/${TAG}_ncu.log 2>&1
  echo "== $1: $2 $3"; python tools/summarize_launches.py gpurun_out/${TAG}_launches_$1.csv 30 | grep -E "total|mix_bwd"; }
prof base X=0 Y=0
prof mb4_72 GLOWK_LIB=$PWD/pytorch_glow_b200/libglowk_mb4.so Y=0
prof mb4_56 GLOWK_LIB=$PWD/pytorch_glow_b200/libglowk_mb4.so GLOWK_MIXBWD_SMEM_KB=56
prof mb4_44 GLOWK_LIB=$PWD/pytorch_glow_b200/libglowk_mb4.so GLOWK_MIXBWD_SMEM_KB=44
