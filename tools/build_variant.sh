#!/bin/bash
# usage: tools/build_variant.sh <name> [-DNAME=VALUE ...]   ->  build/variants/<name>.so (same flags as the Makefile)
# wavefront.cu recompiled with the extra defines and linked against the current objects of csrc/ (run make first).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/cpu-raytracing-rt_b200/csrc
mkdir -p $root/build/variants
cd $src
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden,-Wall -Xptxas -v "$@" -c wavefront.cu -o $root/build/variants/$name.wf.o 2> $root/build/variants/$name.ptxas.log \
  || (cat $root/build/variants/$name.ptxas.log; exit 1)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/build/variants/$name.so scene_build.o scene_cache.o loaders.o fast_build.o capi.o comm.o kernels.o \
  $root/build/variants/$name.wf.o -cudart shared -lpthread -ldl
rm -f $root/build/variants/$name.wf.o
