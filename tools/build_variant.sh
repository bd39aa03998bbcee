#!/bin/bash
# tools/build_variant.sh name -DRT_X=1 ... : builds build_variants/name.so with extra nvcc flags
set -e
mkdir -p build_variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC "$@" -o build_variants/$name.so raytracing.jl_b200/csrc/rt_b200.cu -ldl
