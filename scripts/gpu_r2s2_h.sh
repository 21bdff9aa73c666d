#!/bin/bash
mkdir -p gpurun_out
{
echo "== new build"; timeout 300 python scripts/bench_graph.py --engine --default-only
cp ai-generated-gtav_b200/libgtav_b200.so /tmp/new.so; cp build_tmp/libgtav_b200_head.so ai-generated-gtav_b200/libgtav_b200.so
echo "== HEAD build (counter rendezvous, 48bc941)"; timeout 300 python scripts/bench_graph.py --engine --default-only
cp /tmp/new.so ai-generated-gtav_b200/libgtav_b200.so
echo "== new build again"; timeout 300 python scripts/bench_graph.py --engine --default-only
timeout 900 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "tagged vs|passed|failed|Error|error" | tail -20
} > gpurun_out/r2s2_h.log 2>&1
grep -E "==|last_frame|passed|failed|tagged vs" gpurun_out/r2s2_h.log | cut -c1-250
