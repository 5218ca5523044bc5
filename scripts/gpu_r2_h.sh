#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python scripts/time_small_shard.py 32 2>&1 | grep -E "round|whole|EM iteration|mstep|estep" | head -12
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print(d['value'], d['ms_per_step'], {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'evals/step', d.get('hstep_evals_per_step'), 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'H', d['roofline_hstep']['frac'], 'M', d['roofline_mstep']['frac'], d['roofline_mstep']['ms_per_launch'])
PY
