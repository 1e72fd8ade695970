#!/bin/bash
# Builds scone_b200/libscone_b200.so (CUDA kernels + C ABI + host driver) for sm_100a, in-tree.
# -fmad=false: the reference is built without FMA (gfortran -O3, baseline x86-64); IDs must be bit-exact.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
      -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2,-ffp-contract=off -diag-suppress 39 \
      -shared engine.cu host/physics_package.cpp -o ../libscone_b200.so "$@"
