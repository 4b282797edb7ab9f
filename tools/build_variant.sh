#!/bin/bash
# build an experimental variant of the library: tools/build_variant.sh <name> [extra nvcc flags...]
# -> gpurun_variants/lib_<name>.so (travels with the snapshot; select with DGTTA_LIB_PATH)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
cd dg_tta_b200/csrc
srcs=$(ls *.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC "$@" -o ../../gpurun_variants/lib_${name}.so $srcs
echo built gpurun_variants/lib_${name}.so
