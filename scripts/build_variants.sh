#!/bin/bash
# Timing-experiment builds of the library: scripts/build_variants.sh "<nvcc -D flags>" <tag>  ->  lib_exp/libdllm_b200_<tag>.so
# (select with DLLM_B200_LIB=...; results of some variants are deliberately wrong — never shipped)
set -e
cd "$(dirname "$0")/../diffusion-llm-rs_b200"
mkdir -p lib_exp build_exp
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr"
for f in gemv_mma umma_gemm quant_kernels; do
  /usr/local/cuda/bin/nvcc $FLAGS $1 -c csrc/$f.cu -o build_exp/${f}_$2.o &
done
wait
OBJS=""
for f in api weight_kernels gemv_simt sample_kernels tp; do OBJS="$OBJS build/$f.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib_exp/libdllm_b200_$2.so $OBJS build_exp/gemv_mma_$2.o build_exp/umma_gemm_$2.o build_exp/quant_kernels_$2.o -lnccl
echo built lib_exp/libdllm_b200_$2.so
