#!/bin/sh
# same-box A/B of engine variants (gpurun: build/ is git-ignored but travels to the GPU box).  usage: mk.sh name [nvcc flags]  -> build/ab/name.so
name=$1; shift
cd /root/repo/open-symuvia_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 \
  -Xcompiler -fPIC,-O2,-ffp-contract=off,-fvisibility=default -shared -cudart static \
  -o /root/repo/build/ab/$name.so engine.cu symexp.cpp "$@" 2>&1 | grep -v "^$" | tail -3
echo built $name
