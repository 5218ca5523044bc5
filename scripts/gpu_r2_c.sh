#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -30
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multirank.py 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print(d['value'], d['ms_per_step'], d['split_ms'], 'evals/step', d.get('hstep_evals_per_step'), 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
PY
