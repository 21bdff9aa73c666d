#!/bin/bash
# A/B of a compile-time switch of gemm_skinny.cu on one box: engine step bench with the shipped build, then with -D$1, then
# the shipped build again.   bash scripts/ab_skinny_define.sh GTAV_SK_FENCE_ALL
set -e
cd "$(dirname "$0")/.."
P=ai-generated-gtav_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --cudart shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
run() { timeout 300 python scripts/bench_graph.py --engine 2>&1 | head -1 | cut -c1-220; }
echo "shipped:"; run
cp $P/build/gemm_skinny.o /tmp/gemm_skinny.o.keep
nvcc $FLAGS -D$1 -c $P/csrc/gemm_skinny.cu -o $P/build/gemm_skinny.o
nvcc -shared --cudart shared -o $P/libgtav_b200.so $P/build/*.o -Xlinker -rpath=/usr/local/cuda/lib64
echo "with -D$1:"; run
cp /tmp/gemm_skinny.o.keep $P/build/gemm_skinny.o
nvcc -shared --cudart shared -o $P/libgtav_b200.so $P/build/*.o -Xlinker -rpath=/usr/local/cuda/lib64
echo "shipped again:"; run
