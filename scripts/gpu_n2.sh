#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','steps','warmup') if k in d}, d.get('e2e',{}).get('value'), 'c5', d['c5']['value'], d['c5'].get('per_gpu_frames_per_s'))
PY
