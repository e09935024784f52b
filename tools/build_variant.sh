#!/bin/bash
# Tuning variant of the forward kernels of one plan (PLAN=<n_fft>, default 1024):  [PLAN=2048] tools/build_variant.sh <name> [-D...]
# compiles csrc/stft_fwd_plan.cu (plan 1024, headline variant only unless -DACIDS_FWD_ONLY is overridden) with the extra
# flags and links it with the product's other objects into acids_transforms_b200/variants/<name>.so
# (load it with ACIDS_B200_LIB=<path>).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/acids_transforms_b200
mkdir -p $pkg/variants
only="-DACIDS_FWD_ONLY=2"
for a in "$@"; do case $a in -DACIDS_FWD_ONLY=*) only="";; esac; done
nvcc -std=c++17 -O3 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -DACIDS_FWD_PLAN_N=${PLAN:-1024} $only "$@" -Xptxas -v -c $pkg/csrc/stft_fwd_plan.cu -o $pkg/variants/$name.o 2> $pkg/variants/$name.log
objs=$(ls $pkg/build/*.o | grep -v stft_fwd_plan_${PLAN:-1024}.o)
nvcc -shared -o $pkg/variants/$name.so $objs $pkg/variants/$name.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
grep -E "registers|spill" $pkg/variants/$name.log | paste - - | sed 's/ptxas info    ://g'
