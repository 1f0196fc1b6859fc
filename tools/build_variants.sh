#!/bin/bash
# Builds compile-time variants of libb200geo.so into gpurun_variants/ (git-ignored; travels to the GPU box):
#   tools/build_variants.sh name1 "-DFLAG=1 ..." name2 "..."
set -e
cd "$(dirname "$0")/../ransac_with_homography_b200/csrc"
mkdir -p ../../gpurun_variants
while [ $# -gt 1 ]; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr --extended-lambda $2 -shared -o ../../gpurun_variants/lib_$1.so b200geo.cu &
  shift 2
done
wait
ls -la ../../gpurun_variants/
