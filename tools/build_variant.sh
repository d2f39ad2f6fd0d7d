#!/bin/bash
# Dev tool: build a named variant of libpetmh.so with extra -D flags into build/variants/ (timed by tools/variant_probe.sh).
# Usage: tools/build_variant.sh NAME [-DFLAG=V ...]
set -e
name=$1; shift
mkdir -p build/variants
cd pet_posterior_distribution_b200/csrc
nvcc "$@" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -shared \
  -o ../../build/variants/libpetmh_$name.so petmh.cu 2> ../../build/variants/$name.log
grep -A3 "Compiling entry function.*mh_sweep_kernelILi0ELb0ELi0" ../../build/variants/$name.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $name"
