#!/bin/bash
# A/B builds of libear_b200.so: scripts/build_variant.sh NAME [-DEARB_...=...]...  ->  build_variants/NAME.so
# (selected at run time with EAR_B200_LIB=build_variants/NAME.so; the directory travels to the GPU box, not into git)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-O3,-ffp-contract=off "$@" -shared -o build_variants/$name.so \
  ear_b200/csrc/ear_b200.cu ear_b200/csrc/bvh_build.cpp -lcudart
echo built build_variants/$name.so
