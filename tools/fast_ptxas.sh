#!/bin/bash
# ptxas report (registers, spills) of the production tile only: seconds instead of minutes
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -DBB_FAST_BUILD "$@" -Xptxas -v -cubin -o /tmp/fast.cubin mixed-integer-optimal-control---algorithm-tools_b200/csrc/kernel_wavefront.cu 2>&1 | grep -A2 "ELb0ELi4ELi128" | grep -v "^--"
