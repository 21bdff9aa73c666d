mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1; tail -1 gpurun_out/bench_c2.log | cut -c1-250
timeout 600 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; tail -1 gpurun_out/bench_c3.log | cut -c1-250
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/bench_c5.log 2>&1; tail -1 gpurun_out/bench_c5.log | cut -c1-250
timeout 600 python scripts/bench_vae.py > gpurun_out/bench_vae.log 2>&1; cat gpurun_out/bench_vae.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-120
