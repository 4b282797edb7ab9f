#!/bin/bash
# A/B of library variants on any kernel_times line: tools/ab_any.sh <kernel_times arg> <filter-regex> variantA variantB ...  ("default" = in-tree build)
what=$1; filter=$2; shift; shift
for round in 1 2; do
  for v in "$@"; do
    echo "== $v (round $round)"
    if [ "$v" = default ]; then python tools/kernel_times.py $what 2>&1 | grep -E "$filter"
    else DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_$v.so python tools/kernel_times.py $what 2>&1 | grep -E "$filter"; fi
  done
done
