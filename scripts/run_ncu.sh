# ncu --set full of the weight-streaming GEMM (4 consecutive launches of the real last-frame step: fc1, fc2+LN, qkv(+attn), out+LN)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 300 -c 4 -o gpurun_out/prof_skinny python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-dense > gpurun_out/ncu_skinny.log 2>&1; tail -2 gpurun_out/ncu_skinny.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
