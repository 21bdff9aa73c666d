#!/bin/bash
# One gpurun call: kernel tests, model parity, micro-benchmarks.  Each stage in its own process with a
# timeout so a trapped kernel cannot take the rest down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -${TAILN:-25} gpurun_out/$name.log; }
TAILN=40 run t_gemm python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" --no-header -p no:cacheprovider
TAILN=30 run t_kernels python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not gemm" --no-header -p no:cacheprovider
TAILN=60 run t_model python -m pytest tests/test_model_gpu.py -m gpu -q -s --no-header -p no:cacheprovider
TAILN=40 run b_kernels python scripts/bench_kernels.py
