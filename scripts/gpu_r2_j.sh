#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --config config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_config3.json 2> gpurun_out/r2j_bench_config3.err
tail -3 gpurun_out/r2j_bench_config3.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2j_bench_config3.json'))
    print('config3', d['value'], d['ms_per_step'], {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['roofline']['factor_columns'])
except Exception as e: print('config3 ERR', e)
PY
