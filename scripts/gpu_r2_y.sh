#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -vE "^Iteration|^Trial|Initializ|Fitting|Inferring|Done" | tail -4 | tee gpurun_out/r2y_pytest.log
for v in tmay notmay; do
  if [ $v = notmay ]; then export VLGP_MSTEP_NO_TMA_Y=1; fi
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2y_bench_$v.json 2> gpurun_out/r2y_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2y_bench_$v.json'))
print('$v', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', round(d['e2e']['value'],2), 'M', round(d['roofline_mstep']['frac'],3), round(d['roofline_mstep']['ms_per_launch'],4))
PY
done
