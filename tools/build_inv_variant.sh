#!/bin/bash
# Tuning variant of the inverse kernels:  tools/build_inv_variant.sh <name> [-D...]  -> acids_transforms_b200/variants/<name>.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/acids_transforms_b200
mkdir -p $pkg/variants
nvcc -std=c++17 -O3 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  "$@" -Xptxas -v -c $pkg/csrc/istft.cu -o $pkg/variants/$name.o 2> $pkg/variants/$name.log
objs=$(ls $pkg/build/*.o | grep -v "/istft.o")
nvcc -shared -o $pkg/variants/$name.so $objs $pkg/variants/$name.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
grep -E "Compiling|registers|spill" $pkg/variants/$name.log | paste - - - | sed 's/ptxas info    ://g' | grep "istft_ola_kernelINS_4PlanILi1024" | sed 's/.*EvNS_9InvParamsE//'
