#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print(d['value'], d['ms_per_step'], d['split_ms'], 'evals/step', d.get('hstep_evals_per_step'), 'e2e', d['e2e']['value'])
PY
VLGP_HSTEP_SCIPY=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_scipy.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_scipy.json'))
print('scipy-driven:', d['value'], d['ms_per_step'], d['split_ms'], 'evals/step', d.get('hstep_evals_per_step'))
PY
ncu --set full --clock-control none --import-source on -k regex:estep_seg_kernel -s 1 -c 1 -o gpurun_out/r2_estep_seg_kernel -f python scripts/profile_driver.py 2 > gpurun_out/r2_ncu_estep_seg_kernel.log 2>&1
tail -3 gpurun_out/r2_ncu_estep_seg_kernel.log
