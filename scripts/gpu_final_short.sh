#!/bin/bash
# Re-verification after a small change: GPU tests, smoke(), a short bench line (kernel legs only).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_short_gputests.log 2>&1; echo "pytest rc $?"; tail -1 gpurun_out/final_short_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_short_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/final_short_smoke.log | cut -c1-200
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-dense --no-c5 --no-c3 --no-eager --no-cpu-baseline > gpurun_out/final_short_bench.json 2> gpurun_out/final_short_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_short_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_dit_step') if k in d}, d.get('e2e',{}).get('value'), d['roofline']['frac'], d['roofline']['us_per_launch'])
PY
