#!/bin/bash
# Build a variant of libpn2b200.so with extra -D flags for ONE source file (A/B timing of kernel variants on the GPU box):
#   tools/dev/build_variant.sh <name> <source.cu> -DFOO=1 ...   ->  hotrack_b200/variants/<name>/libpn2b200.so
# The other objects are taken from hotrack_b200/build (run python hotrack_b200/build.py first).
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/../.." && pwd)
out=$root/hotrack_b200/variants/$name
mkdir -p $out
base=$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $root/hotrack_b200/csrc/$base.cu -o $out/$base.o
objs=$(ls $root/hotrack_b200/build/*.o | grep -v "/$base.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libpn2b200.so $objs $out/$base.o -lcudart
echo $out/libpn2b200.so
