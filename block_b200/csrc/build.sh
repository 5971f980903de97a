#!/bin/bash
# Builds block_b200/lib/libblockb200.so for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/../build"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function $B2D_EXTRA_NVCC_FLAGS"
$NVCC $FLAGS -c "$HERE/kernels.cu" -o "$HERE/../build/kernels.o" &
$NVCC $FLAGS -x cu -c "$HERE/ctx.cpp" -o "$HERE/../build/ctx.o" &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libblockb200.so" "$HERE/../build/kernels.o" "$HERE/../build/ctx.o" -ldl
echo "built $OUT/libblockb200.so"
# C++ host mirror of the reference's entry points (block_b200/host) + its test driver
CXX=${CXX:-g++}
$CXX -std=c++17 -O2 -Wall -fPIC -shared -o "$OUT/libb2dhost.so" "$HERE/../host/b2d_host.cpp" -L"$OUT" -lblockb200 -Wl,-rpath,'$ORIGIN'
$CXX -std=c++17 -O2 -Wall -o "$OUT/host_mirror_test" "$HERE/../../tests/cpp/host_mirror_test.cpp" -L"$OUT" -lb2dhost -lblockb200 -Wl,-rpath,'$ORIGIN'
echo "built $OUT/libb2dhost.so and $OUT/host_mirror_test"
