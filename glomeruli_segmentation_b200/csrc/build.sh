#!/bin/bash
# Builds libespnet_b200.so in-tree (sm_100a only).  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
      -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
      -o "${OUT:-libespnet_b200.so}" espnet_api.cu "$@"
