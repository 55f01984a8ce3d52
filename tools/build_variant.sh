#!/bin/bash
# usage: tools/build_variant.sh <tag> <extra nvcc flags...>  -> ode_b200/variants/<tag>/libode_b200_single.so (experiments only)
set -e
cd "$(dirname "$0")/../ode_b200/csrc"
tag=$1; shift
mkdir -p ../variants/$tag
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 \
  -Xcompiler -fPIC -Xlinker -Bsymbolic-functions -Xcompiler -Wno-unused-function -shared -cudart shared -ldl "$@" -o ../variants/$tag/libode_b200_single.so odeb_kernels.cu
