#!/bin/bash
# usage: tools/build_variant.sh NAME [-D...]   -> optimalinterpolation_b200/csrc/liboi_exp_NAME.so (experimental builds, OI_LIB=...)
set -e
cd "$(dirname "$0")/../optimalinterpolation_b200/csrc"
name=$1; shift
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $ARCH -fmad=false -Xptxas -v "$@" -c oi_kernels.cu -o /tmp/k_$name.o 2> /tmp/ptxas_$name.log
nvcc -O3 -std=c++17 -Xcompiler -fPIC $ARCH "$@" -c oi_api.cu -o /tmp/a_$name.o
nvcc -shared $ARCH -o liboi_exp_$name.so /tmp/k_$name.o /tmp/a_$name.o -lcudart
echo built $name
