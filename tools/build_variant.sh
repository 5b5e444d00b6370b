#!/bin/bash
# usage: tools/build_variant.sh <name> [-D...]   ->  audio-classification_b200/variants/lib_<name>.so  (kernel A/B experiments)
set -e
name=$1; shift
cd "$(dirname "$0")/../audio-classification_b200"
mkdir -p variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC "$@" -o variants/lib_$name.so csrc/b200sep.cu
echo built variants/lib_$name.so
