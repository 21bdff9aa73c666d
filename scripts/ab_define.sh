#!/bin/bash
# A/B of compile-time switches on one box: rebuild ONE source of the library with extra -D flags, run a command against that
# build, restore the shipped object and run the command again.
#   bash scripts/ab_define.sh gemm_sm100_2cta "-DGTAV_G2_KC=1 -DGTAV_G2_STAGES=7" "python scripts/bench_2cta.py"
cd "$(dirname "$0")/.."
P=ai-generated-gtav_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --cudart shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
link() { nvcc -shared --cudart shared -o $P/libgtav_b200.so $P/build/*.o -Xlinker -rpath=/usr/local/cuda/lib64 2>/dev/null; }
cp $P/build/$1.o /tmp/$1.o.keep
nvcc $FLAGS $2 -c $P/csrc/$1.cu -o $P/build/$1.o 2>/dev/null && link
echo "== with $2:"; eval "$3"
cp /tmp/$1.o.keep $P/build/$1.o; link
echo "== shipped build:"; eval "$3"
