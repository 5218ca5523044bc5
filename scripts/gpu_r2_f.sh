#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/time_estep.py 2>&1 | tail -9
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['value'], d['ms_per_step'], d['split_ms'], 'evals/step', d.get('hstep_evals_per_step'), 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'H', d['roofline_hstep']['frac'], 'M', d['roofline_mstep']['frac'], d['roofline_mstep']['ms_per_launch'])
PY
