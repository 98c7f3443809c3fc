#!/bin/bash
# Round-2 profiling pass, ONE gpurun call:  gpurun --timeout 1500 -- 'bash scripts/r2_gpu_profile.sh [tag]'
#  1. ncu launch list (gpu__time_duration.sum, --clock-control none) of a reduced bench command
#  2. ncu --set full captures of the search kernels and the mb-tree walk (single 1080p stream, scripts/profile_la.py)
# Results land in gpurun_out/ ; scripts/summarize_profiles.py r2 turns them into profiles/*.txt and profiles/ncu_metrics_r2.json
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --launch-skip 4000 -c 4000 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --streams 2 --frames-per-step 24 --clip-frames 60 --no-e2e --no-cpu-baseline --no-worst-case > gpurun_out/launches_bench_$TAG.log 2>&1
for K in me_pass_kernel me_verify_kernel tree_chain_kernel frontend_kernel intra_kernel finalize_kernel; do
  SKIP=24; CNT=3
  [ $K = tree_chain_kernel ] && SKIP=3 && CNT=1
  [ $K = me_verify_kernel ] && CNT=2
  [ $K = frontend_kernel ] && SKIP=8 && CNT=1
  [ $K = intra_kernel ] && SKIP=8 && CNT=1
  [ $K = finalize_kernel ] && SKIP=20 && CNT=1
  RC_LOOKAHEAD=40 timeout 300 $NCU --set full --import-source on -k regex:$K --launch-skip $SKIP -c $CNT -f -o gpurun_out/${K}_$TAG \
      python scripts/profile_la.py 60 > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out | tail -20
