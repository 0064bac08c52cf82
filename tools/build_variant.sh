#!/bin/bash
# Development aid: build/libmpb200_<tag>.so from the same sources with extra -D flags, for A/B timing
# on the GPU box through MPB200_LIBRARY.   usage: tools/build_variant.sh tag [-DNAME=VALUE ...]
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p build
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
[ build/fftconv.o -nt matching-pursuit_b200/csrc/fftconv.cu ] || nvcc $F -c matching-pursuit_b200/csrc/fftconv.cu -o build/fftconv.o
[ build/gemm_corr.o -nt matching-pursuit_b200/csrc/gemm_corr.cu ] || nvcc $F -c matching-pursuit_b200/csrc/gemm_corr.cu -o build/gemm_corr.o
nvcc $F "$@" -c matching-pursuit_b200/csrc/mpb200.cu -o build/mpb200_$tag.o
nvcc -shared -o build/libmpb200_$tag.so build/mpb200_$tag.o build/fftconv.o build/gemm_corr.o -lcudart
echo build/libmpb200_$tag.so
