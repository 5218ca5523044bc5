#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hstep or vem_three or fit_tutorial or omega_traj" 2>&1 | grep -vE "^Iteration|^Trial|Initializ|Fitting|Inferring|Done|^[0-9.]+s$" | tail -8
for v in schur noschur; do
  if [ $v = noschur ]; then export VLGP_HSTEP_NO_SCHUR=1; fi
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2s_bench_$v.json 2> gpurun_out/r2s_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2s_bench_$v.json'))
print('$v', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', round(d['e2e']['value'],2), 'H', round(d['roofline_hstep']['frac'],3), round(d['roofline_hstep']['ms_total'],3), d['roofline_hstep']['launches_timed'], d['roofline_hstep']['evaluations'])
PY
done
