#!/bin/sh
# Builds libqgt_b200.so in-tree for sm_100a (the same command __graft_entry__.build() runs).
set -e
cd "$(dirname "$0")/quantum_geometric_tensor_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" \
     -shared -o ../libqgt_b200.so kernels.cu capi.cu dist.cu plan.cpp natgrad.cpp -ldl -lpthread
