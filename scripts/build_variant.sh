#!/bin/bash
# Builds a variant of liblsi.so that differs in the flags of ONE translation unit (kernel experiments):
#   scripts/build_variant.sh <name> <source.cu> [extra nvcc flags...]
# -> integrator-benchmark_b200/lib/variants/liblsi_<name>.so, to be used with LSI_LIB_PATH=...
# The other objects come from the regular build (python integrator-benchmark_b200/build.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/integrator-benchmark_b200
name=$1; src=$2; shift 2
mkdir -p $PKG/lib/variants
base=$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O3 \
     --expt-relaxed-constexpr "$@" -x cu -c $PKG/csrc/$src -o $PKG/lib/variants/${base}_$name.o
objs=""
for o in $PKG/lib/*.o; do
  if [ "$(basename $o .o)" == "$base" ]; then objs="$objs $PKG/lib/variants/${base}_$name.o"; else objs="$objs $o"; fi
done
nvcc -shared -o $PKG/lib/variants/liblsi_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcudart
echo $PKG/lib/variants/liblsi_$name.so
