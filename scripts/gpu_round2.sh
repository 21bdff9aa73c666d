#!/bin/bash
# Round-2 profiler captures in one gpurun call (logs and reports land in gpurun_out/):
#   launch list of a config-1-shaped rollout (shares of a step), ncu --set full of four consecutive weight-streaming GEMM
#   launches of the real last-frame step, and of the tcgen05 attention kernel inside the VAE decode.
mkdir -p gpurun_out
QUIET="--no-cpu-baseline --no-dense --no-c5 --no-c3 --no-c1 --no-eager"
for st in ${STAGES:-list skinny attn}; do
case $st in
list) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file gpurun_out/r02_launches_c1.csv python bench.py --workload c1 --steps 1 --warmup 1 $QUIET > gpurun_out/r02_ncu_list.log 2>&1; tail -1 gpurun_out/r02_ncu_list.log | cut -c1-160 ;;
skinny) timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 300 -c 4 -f -o gpurun_out/r02_prof_skinny python bench.py --workload c1 --steps 1 --warmup 1 $QUIET > gpurun_out/r02_ncu_skinny.log 2>&1; tail -1 gpurun_out/r02_ncu_skinny.log | cut -c1-160 ;;
attn) timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 1 -f -o gpurun_out/r02_prof_attn_tc python scripts/run_attn_tc.py 576 32 > gpurun_out/r02_ncu_attn.log 2>&1; tail -1 gpurun_out/r02_ncu_attn.log | cut -c1-160 ;;
esac
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c1.csv
