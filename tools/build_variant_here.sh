#!/bin/bash
# usage: tools/build_variant_here.sh <name> <extra nvcc flags...>  -> mfcc-rust_b200/lib/variants/<name>.so  (fast512c.cu only)
set -e
cd "$(dirname "$0")/../mfcc-rust_b200"
name=$1; shift
mkdir -p lib/variants build/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
  -Xcompiler -fPIC,-fvisibility=hidden,-Wall -I../include -Icsrc "$@" -c csrc/ss_kernels_fast512c.cu -o build/variants/$name.o
objs=$(ls build/*.o | grep -v fast512c)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/$name.so $objs build/variants/$name.o -lpthread
echo built lib/variants/$name.so
