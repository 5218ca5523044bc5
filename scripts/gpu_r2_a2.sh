#!/bin/bash
mkdir -p gpurun_out
python scripts/time_fit.py 20 2>&1 | head -2
for v in horner estrin; do
  if [ $v = estrin ]; then export VLGP_MSTEP_ESTRIN=1; fi
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2a2_bench_$v.json 2> gpurun_out/r2a2_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2a2_bench_$v.json'))
print('$v', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', round(d['e2e']['value'],2), 'M', round(d['roofline_mstep']['frac'],3), round(d['roofline_mstep']['ms_per_launch'],4))
PY
done
VLGP_MSTEP_ESTRIN=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mstep or vem_three or fit_tutorial" 2>&1 | tail -2
