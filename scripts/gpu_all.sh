#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_fit.py 20 2>&1 | head -4
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['value'], d['ms_per_step'], {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'H', d['roofline_hstep']['frac'], 'M', d['roofline_mstep']['frac'], d['roofline_mstep']['ms_per_launch'])
PY
