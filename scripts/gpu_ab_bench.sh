#!/bin/bash
for rep in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_fused.json 2>/dev/null
VLGP_ESTEP_NO_FUSED=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_unfused.json 2>/dev/null
python - <<'PY'
import json
for f in ('ab_fused','ab_unfused'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, round(d['value'],2), round(d['ms_per_step'],2), 'estep', round(d['roofline']['ms_per_launch'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, d['clocks'])
PY
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
