#!/bin/bash
# Build a kernel-variant copy of the library for A/B runs (PANIB200_LIB=tools/variants/NAME.so):
#   bash tools/build_variant.sh NAME -DPANIB_K1_WARP=1 -DPANIB_K1_THREADS=32 ...
set -e
NAME=$1; shift
cd "$(dirname "$0")/../pyani_plus_b200/csrc"
mkdir -p ../../tools/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared \
  "$@" -Xptxas -v -o ../../tools/variants/$NAME.so api.cu sketch.cu pairwise.cu index.cu synth.cu hostpack.cpp hostio.cpp 2>&1 \
  | grep -A1 "sketch_hash_kernelILi31" | grep -E "registers|spill" | head -4
