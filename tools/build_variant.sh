#!/bin/bash
# A/B builds of libcamus_b200.so: tools/build_variant.sh <name> <extra nvcc flags...>  -> build_variants/libcamus_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall --expt-relaxed-constexpr \
  -I include -shared -o build_variants/libcamus_${name}.so camus_b200/csrc/camus_b200.cu -cudart static -ldl "$@"
echo built build_variants/libcamus_${name}.so
