#!/bin/bash
# dev helper: build a variant of the fused kernel file with extra nvcc flags into build/ab/NAME.so
# usage: tools/build_variant.sh NAME [-DFLAG=..]...
set -e
cd "$(dirname "$0")/../zhusuan-pytorch_b200"
name=$1; shift
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static \
  -I../include "$@" -c csrc/zs_fused_iw.cu -o build/ab/$name.o 2>&1 | grep -v "1886-D\|extern __attribute__\|\^\|^$\|Remark" || true
objs=$(ls build/*.o | grep -v zs_fused_iw.o)
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o build/ab/$name.so build/ab/$name.o $objs
echo built build/ab/$name.so
