#!/bin/bash
# Ablation of the first layers (stem, pixel-pair Cin = 32 layers) with the profiling build's Y3_DBG knobs:
# 1 epilogue drains TMEM only, 2 no global stores, 4 no A loads (stem: no image loads), 8 no MMAs, 32 stem producers skip
# their shared-memory stores.  Per-kernel ncu times of the first 8 conv launches of a forward pass.
# usage: bash tools/first_layers_ablation.sh <tag>
TAG=${1:-abl}
for d in ${Y3_ABL_SET:-0 1 2 4 8 32 3}; do
  env Y3_PROF_LIB=1 Y3_DBG=$d Y3_CHAIN=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_' -s 75 -c 8 --csv \
      --log-file gpurun_out/${TAG}_dbg$d.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --ncu > gpurun_out/${TAG}_dbg$d.log 2>&1
  echo "Y3_DBG=$d: $(grep gpu__time gpurun_out/${TAG}_dbg$d.csv | awk -F'","' '{gsub(/"/,"",$NF); printf "%s=%.0f ", substr($5,6,22), $NF/1000}')"
done
