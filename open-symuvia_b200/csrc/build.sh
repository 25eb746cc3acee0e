#!/bin/sh
# Builds libSymuViaB200.so (sm_100a only; -fmad=false keeps fp64 arithmetic contraction-free like the reference's x86 build).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 \
  -Xcompiler -fPIC,-O2,-ffp-contract=off,-fvisibility=default -shared -cudart static \
  -o ../libSymuViaB200.so engine.cu symexp.cpp "$@"

# host-only synthetic network generator used by bench.py
${CXX:-g++} -O2 -std=c++17 -fPIC -shared -o ../libsymgen.so gen_grid.cpp
