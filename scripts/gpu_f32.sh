#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float32 or estep_golden" 2>&1 | tail -15
for cfg in config2 config3; do
python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu --dtype f32 > gpurun_out/r2n_bench_${cfg}_f32.json 2> gpurun_out/r2n_bench_${cfg}_f32.err
tail -3 gpurun_out/r2n_bench_${cfg}_f32.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2n_bench_${cfg}_f32.json'))
print('$cfg f32', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'estep ms/launch', round(d['roofline']['ms_per_launch'],3), d['roofline']['launches_timed'])
PY
done
python bench.py --config config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2n_bench_config3_f64.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r2n_bench_config3_f64.json'))
print('config3 f64', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'estep ms/launch', round(d['roofline']['ms_per_launch'],3), d['roofline']['launches_timed'])
PY
