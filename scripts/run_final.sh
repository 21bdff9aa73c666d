mkdir -p gpurun_out
# STAGES: c2 c3 c5 vae ncu
for st in ${STAGES:-c2 c3 ncu}; do
case $st in
c2) timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1; tail -1 gpurun_out/bench_c2.log | cut -c1-400 ;;
c3) timeout 600 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; tail -1 gpurun_out/bench_c3.log | cut -c1-300 ;;
c5) timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/bench_c5.log 2>&1; tail -1 gpurun_out/bench_c5.log | cut -c1-300 ;;
vae) timeout 600 python scripts/bench_vae.py > gpurun_out/bench_vae.log 2>&1; cat gpurun_out/bench_vae.log ;;
ncu) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-120 ;;
ref) timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300 ;;
esac
done
