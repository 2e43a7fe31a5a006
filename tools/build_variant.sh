#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags]  -> tools/variants/lib_<name>.so (git-ignored; travels to the GPU box)
set -e
name=$1; shift
mkdir -p tools/variants
cd membranealefem.jl_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o ../../tools/variants/lib_$name.so maf_api.cu
