#!/bin/bash
# quick A/B build of the rollout kernels: tools/abbuild.sh <name> [extra nvcc flags...]  ->  ab_build/<name>.so
name=$1; shift
mkdir -p ab_build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -prec-div=false \
     -prec-sqrt=false -ftz=true -DPPR_AB_ONLY "$@" -o ab_build/$name.so ppr_diffphys_b200/csrc/ppr_kernels.cu
