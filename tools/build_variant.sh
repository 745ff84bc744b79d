#!/bin/bash
# Build a kernel-variant copy of the library: tools/build_variant.sh TAG [-DMACRO=VALUE ...]
# -> maniac-mc.github.io_b200/variants/libmaniac_gpu_TAG.so   (select with MANIAC_GPU_LIB=<path>)
set -e
TAG=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/maniac-mc.github.io_b200/csrc
mkdir -p $ROOT/maniac-mc.github.io_b200/variants
cd $CSRC
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o $ROOT/maniac-mc.github.io_b200/variants/libmaniac_gpu_$TAG.so.tmp maniac_gpu.cu mgpu_nccl.cu maniac_host.cpp mgpu_table.cpp -ldl -lpthread
mv $ROOT/maniac-mc.github.io_b200/variants/libmaniac_gpu_$TAG.so.tmp $ROOT/maniac-mc.github.io_b200/variants/libmaniac_gpu_$TAG.so
echo built $TAG
