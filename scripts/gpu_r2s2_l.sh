#!/bin/bash
mkdir -p gpurun_out
{
echo "== dealloc after the drain"; timeout 300 python scripts/bench_graph.py --engine --default-only
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== trace"; timeout 300 python scripts/trace_step.py 8 9
} > gpurun_out/r2s2_l.log 2>&1
grep -E "==|last_frame|passed|failed|#" gpurun_out/r2s2_l.log | cut -c1-215
