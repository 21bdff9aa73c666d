TMO=400 STAGES="tests_k" bash scripts/gpu_round.sh
timeout 400 python scripts/bench_graph.py 2>&1 | grep -E "gemm_skinny|step body|# " | head -30
timeout 600 python bench.py --no-cpu-baseline --steps 1 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c2_v3.json; cut -c1-300 gpurun_out/bench_c2_v3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 300 -c 4 -o gpurun_out/prof_skinny python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/ncu_skinny.log 2>&1; tail -2 gpurun_out/ncu_skinny.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 40 -c 3 -o gpurun_out/prof_tiled python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --algorithm dense > gpurun_out/ncu_tiled.log 2>&1; tail -2 gpurun_out/ncu_tiled.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
