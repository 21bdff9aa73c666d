#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list.  Each stage in its own process with a
# timeout so a trapped kernel cannot take the rest down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -${TAILN:-25} gpurun_out/$name.log; }
STAGES=${STAGES:-"tests smoke bench ncu"}
for st in $STAGES; do
case $st in
tests) TAILN=30 run t_all python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider ;;
tests_mega) TAILN=25 run t_mega python -m pytest tests/test_model_gpu.py -m gpu -q -s --no-header -p no:cacheprovider -k "last_frame or step_kernel" ;;
tests_k) TAILN=8 run t_kernels python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "skinny or temporal or frame" ;;
smoke) TAILN=5 run smoke python -c "import __graft_entry__ as g; g.smoke()" ;;
bench) TAILN=5 run bench python bench.py ;;
bench_dense) TAILN=5 run bench_dense python bench.py --algorithm dense --steps 1 --warmup 3 --no-cpu-baseline ;;
bench_c3) TAILN=5 run bench_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline ;;
bench_c1) TAILN=5 run bench_c1 python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline ;;
kbench) TAILN=40 run b_kernels python scripts/bench_kernels.py ;;
engine) TAILN=12 run b_engine python scripts/bench_graph.py --engine ;;
klight) TAILN=60 run b_light python scripts/bench_kernels.py --light ;;
ncu) TAILN=3 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-dense ;;
ncu_full) TAILN=3 run ncu_full ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-gemm_skinny} -s ${NCU_S:-200} -c 4 -o gpurun_out/prof_${NCU_NAME:-skinny} python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-dense ;;
esac
done
