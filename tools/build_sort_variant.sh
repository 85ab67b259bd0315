#!/bin/bash
# Experiment variant of libb200arrays.so: only the 4-byte-key sort object is recompiled with the given -D flags.
# usage: tools/build_sort_variant.sh NAME -DFOO=1 ...
set -e
cd "$(dirname "$0")/../cuda.jl_b200/csrc"
name=$1; shift
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
$NVCC $FLAGS "$@" -DB200_KEY_BYTES=4 -c sort_inst.cu -o /tmp/sort_k4_$name.o
others=$(ls build/*.o | grep -v "/sort_k4.o")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libb200arrays_$name.so /tmp/sort_k4_$name.o $others
rm -f /tmp/sort_k4_$name.o
echo built ../libb200arrays_$name.so
