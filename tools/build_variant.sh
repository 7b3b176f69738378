#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]   -> build/libbb_<name>.so (A/B builds of the library; select with $BELLMAN_B200_LIB)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
C=mixed-integer-optimal-control---algorithm-tools_b200/csrc
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC -shared -ldl "$@" \
  -o build/libbb_$NAME.so $C/bb200_api.cu $C/bb200_multi.cu $C/kernels_common.cu $C/kernel_wavefront.cu $C/kernel_stage_pruned.cu $C/kernels_microbench.cu
echo built build/libbb_$NAME.so
