#!/bin/bash
# Round-2 closing single-GPU record at HEAD: GPU suite, smoke(), both bench arms at the driver's settings, config 4 on one
# GPU (the denominator of the 8-GPU config-4 line), config 3, whole fit / full-trial inference, launch list of the same
# command as the bench step, ncu --set full of the M-step statistics kernel (new this session) and the E-step kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -vE "^Iteration|^Trial|Initializ|Fitting|Inferring|Done" | tail -4 | tee gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err ) 2>&1 | grep real
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
if [ -z "$SKIP_CONFIGS" ]; then
python bench.py --config config4 --steps 10 --warmup 5 --no-cpu > gpurun_out/r2z_config4_1gpu.json 2>/dev/null
python bench.py --config config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2z_config3_f64.json 2>/dev/null
python bench.py --config config3 --steps 5 --warmup 3 --no-cpu --dtype f32 > gpurun_out/r2z_config3_f32.json 2>/dev/null
fi
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print('ours', round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])
print('roofline', round(d['roofline']['frac'],3), round(d['roofline']['ms_per_launch'],3), d['roofline']['factor_columns'], 'H', round(d['roofline_hstep']['frac'],3), 'M', round(d['roofline_mstep']['frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'))
r=json.load(open('gpurun_out/r2z_ref.json'))
print('ref', r['value'], r['cpu_baseline']['sample'][:120])
for f in ('config4_1gpu','config3_f64','config3_f32'):
    d=json.load(open('gpurun_out/r2z_%s.json'%f))
    print(f, round(d['value'],2), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'e2e', round(d['e2e']['value'],2))
PY
python scripts/time_fit.py 20 2>&1 | head -2
python scripts/time_infer.py config2 20 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches.csv python scripts/profile_driver.py 8 > /dev/null 2>&1
for ks in mstep_stats_tma_kernel:30 estep_seg_kernel:5 hstep_segment_schur_kernel:20; do
  k=${ks%%:*}; skip=${ks##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2z_$k -f python scripts/profile_driver.py 7 > gpurun_out/r2z_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2z_ncu_$k.log
done
ls gpurun_out | grep r2z | tr '\n' ' '
