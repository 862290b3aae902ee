#!/bin/bash
# usage: tools/build_variant.sh <tag> [extra nvcc flags...]   -> build/variants/libblsgpu_<tag>.so
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  "$@" -o build/variants/libblsgpu_$tag.so nim_blscurve_b200/csrc/blsgpu.cu
echo "built $tag"
