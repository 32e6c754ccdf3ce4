#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=...]: an alternative build of the library for kernel tuning
# (loaded with CLB_LIBRARY_PATH=build/exp/NAME.so); build/ is git-ignored but travels with gpurun.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/exp
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  "$@" -o build/exp/$name.so climaland.jl_b200/csrc/clb_api.cu -ldl
echo build/exp/$name.so
