#!/bin/bash
# Builds libllkv_gpu.so (sm_100a) next to the sources.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p build
$NVCC $FLAGS -c scan_kernel.cu -o build/scan_kernel.o &
$NVCC $FLAGS -c llkv_gpu.cu -o build/llkv_gpu.o &
$NVCC $FLAGS -c fast_kernel.cu -o build/fast_kernel.o &
g++ -O2 -std=c++17 -fPIC -Wall -Wno-nonnull -c compiler.cpp -o build/compiler.o &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libllkv_gpu.so build/scan_kernel.o build/fast_kernel.o build/llkv_gpu.o build/compiler.o -cudart static -ldl -lpthread -lrt
echo built $(pwd)/libllkv_gpu.so
