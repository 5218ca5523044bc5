#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -vE "^Iteration|^Trial|Initializ|Fitting|Inferring|Done" | tail -6 | tee gpurun_out/r2v_pytest.log
python scripts/time_e2e.py 2>&1 | grep -E "pull|vem\(\)|Session" 
for v in 1 0; do
  VLGP_PREFETCH=$v python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2v_bench_pf$v.json 2> gpurun_out/r2v_bench_pf$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2v_bench_pf$v.json'))
print('prefetch=$v', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],2), d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'])
PY
done
