#!/bin/bash
# Developer build of libbpltv with cycle stamps inside grad_factor_kernel (-DBPLTV_FACTOR_TIMING), then
# tools/factor_timing.py on the GPU.  Run `make -j -C bpldenoising_b200/csrc` first (reuses its tblock objects).
set -e
cd "$(dirname "$0")/../bpldenoising_b200/csrc"
mkdir -p ../../tools/_timing
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
     --expt-relaxed-constexpr -DBPLTV_FACTOR_TIMING -c -o /tmp/bpltv_api_timing.o bpltv_api.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_timing/libbpltv_timing.so /tmp/bpltv_api_timing.o obj/tblock_*.o
echo "built tools/_timing/libbpltv_timing.so; on a B200: python tools/factor_timing.py"
