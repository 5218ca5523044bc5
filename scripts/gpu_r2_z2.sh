#!/bin/bash
# closing records without the ncu --set full captures (their reports exceed the 64 MiB that travel back in one call)
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches.csv python scripts/profile_driver.py 8 > /dev/null 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print('ours', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'roofline', round(d['roofline']['frac'],3), 'H', round(d['roofline_hstep']['frac'],3), 'M', round(d['roofline_mstep']['frac'],3))
r=json.load(open('gpurun_out/r2z_ref.json'))
print('ref', r['value'])
PY
du -sh gpurun_out
