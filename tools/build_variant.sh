#!/bin/sh
# build_variant.sh NAME [nvcc flags...] -- experiment build of libb381.so into _variants/lib_NAME.so (tools/quick_bench.py).
# kernels.cu is compiled with the given extra flags; the G1 translation unit is built once and shared.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
C=plonky2-bls12-381-pairing_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p _variants
[ -f _variants/g1_kernels.o ] || nvcc $FLAGS -split-compile 0 -c -o _variants/g1_kernels.o $C/g1_kernels.cu
nvcc $FLAGS "$@" -c -o _variants/kernels_$NAME.o $C/kernels.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _variants/lib_$NAME.so _variants/kernels_$NAME.o _variants/g1_kernels.o
rm -f _variants/kernels_$NAME.o
echo built _variants/lib_$NAME.so
