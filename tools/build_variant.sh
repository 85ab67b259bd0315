#!/bin/bash
# Builds an experiment variant of libb200arrays.so: only the scan objects are recompiled with the given
# -D flags, everything else is linked from the main build.  usage: tools/build_variant.sh NAME -DFOO=1 ...
set -e
cd "$(dirname "$0")/../cuda.jl_b200/csrc"
name=$1; shift
mkdir -p build_$name
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
pairs=$(sed -n 's/^X(\([A-Z0-9]*\), *\([A-Z0-9]*\)).*/\1_\2/p' scan_pairs.def)
only=${ONLY:-$pairs}
for pr in $pairs; do
  in=${pr%_*}; out=${pr#*_}
  if echo " $only " | grep -q " $pr "; then
    $NVCC $FLAGS "$@" -DB200_PAIR_IN=B200_$in -DB200_PAIR_OUT=B200_$out -DB200_PAIR_NAME=scan_pair_$pr -c scan_inst.cu -o build_$name/scan_$pr.o &
  else
    cp build/scan_$pr.o build_$name/scan_$pr.o
  fi
done
wait
others=$(ls build/*.o | grep -v "/scan_[A-Z]")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libb200arrays_$name.so build_$name/scan_*.o $others
rm -rf build_$name
echo built ../libb200arrays_$name.so
