#!/bin/bash
# A/B timing on the GPU box: tools/ab.sh <filter-regex> variantA variantB ...   (interleaved, 2 rounds; "default" = the in-tree build)
filter=$1; shift
for round in 1 2; do
  for v in "$@"; do
    echo "== $v (round $round)  $(nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader)"
    if [ "$v" = default ]; then python tools/kernel_times.py mind 2>&1 | grep -E "$filter"
    else DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_$v.so python tools/kernel_times.py mind 2>&1 | grep -E "$filter"; fi
  done
done
