#!/bin/bash
# Rebuild attn_tc.cu with -D$1 on the box and run the attention tests against that build, then restore the shipped object.
#   bash scripts/ab_attn_define.sh GTAV_ATTN_SINGLE_MREADY
cd "$(dirname "$0")/.."
P=ai-generated-gtav_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --cudart shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
cp $P/build/attn_tc.o /tmp/attn_tc.o.keep
nvcc $FLAGS -D$1 -c $P/csrc/attn_tc.cu -o $P/build/attn_tc.o 2>/dev/null
nvcc -shared --cudart shared -o $P/libgtav_b200.so $P/build/*.o -Xlinker -rpath=/usr/local/cuda/lib64 2>/dev/null
echo "with -D$1:"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention_seq" 2>&1 | tail -8
cp /tmp/attn_tc.o.keep $P/build/attn_tc.o
nvcc -shared --cudart shared -o $P/libgtav_b200.so $P/build/*.o -Xlinker -rpath=/usr/local/cuda/lib64 2>/dev/null
echo "shipped build:"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention_seq" 2>&1 | tail -2
