#!/bin/bash
# A/B timing on the GPU box: tools/ab.sh <filter-regex> variantA variantB ...   (interleaved, 2 rounds)
filter=$1; shift
for round in 1 2; do
  for v in "$@"; do
    echo "== $v (round $round)"
    DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_$v.so python tools/kernel_times.py 2>&1 | grep -E "$filter"
  done
done
