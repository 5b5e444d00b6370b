#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]  -> audio-classification_b200/variants/lib_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p audio-classification_b200/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC "$@" \
  -o audio-classification_b200/variants/lib_$name.so audio-classification_b200/csrc/b200sep.cu
echo built $name
