#!/bin/bash
mkdir -p gpurun_out
STAGES="list skinny" bash scripts/gpu_round2.sh > gpurun_out/r2s2_k_ncu.log 2>&1
{
echo "== out 16 splits"; GTAV_SK_SPLITS=0,16,0,0 timeout 300 python scripts/bench_graph.py --engine --default-only
echo "== out 8 splits"; GTAV_SK_SPLITS=0,8,0,0 timeout 300 python scripts/bench_graph.py --engine --default-only
echo "== default"; timeout 300 python scripts/bench_graph.py --engine --default-only
echo "== unfused"; GTAV_FUSE=0 timeout 300 python scripts/bench_graph.py --engine --default-only
} > gpurun_out/r2s2_k.log 2>&1
tail -5 gpurun_out/r2s2_k_ncu.log; grep -E "==|last_frame" gpurun_out/r2s2_k.log | cut -c1-200
