#!/bin/bash
# Final verification of a build in one gpurun call: GPU tests (parity numbers printed), smoke(), the bench line at the driver's flags.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/final_gputests.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/final_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/final_smoke.log | cut -c1-300
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches') if k in d}, d.get('e2e',{}).get('value'), d.get('clocks'))
r=dict(d.get('roofline')); r.pop('note',None); print('roofline', r)
print('c5', d['c5']['value'], 'c3', d['c3']['value'], d['c3']['ms_per_dit_step'], 'dense', d['dense']['ms_per_dit_step'], d['dense']['frac'])
print('cpu', d['cpu_baseline']['value'], 'c1', d['c1']['product_wall_s'], d['c1']['psnr_db_vs_cpu_fp32'], 'ms_per_dit_step', d.get('ms_per_dit_step'))
PY
