#!/bin/bash
# builds a tuning variant of the library: tools/build_variant.sh <name> <extra nvcc -D flags...>
set -e
name=$1; shift
out=build/variants/$name; mkdir -p $out
for f in api bvh_build render output; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c cutrace_b200/csrc/$f.cu -o $out/$f.o &
done
wait
mkdir -p cutrace_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o cutrace_b200/lib/variants/libcutrace_b200_$name.so $out/*.o -cudart shared
echo built $name
