#!/bin/bash
mkdir -p gpurun_out
n=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus $n --config config5 --steps 6 --warmup 5 --no-cpu > gpurun_out/r2z_bench_8gpu_config5.json 2> gpurun_out/r2z_bench_8gpu_config5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench_8gpu_config5.json'))
print('config5', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'E', round(d['roofline']['ms_per_launch'],3), 'e2e', round(d['e2e']['value'],1), {k: (float('%.2g' % v) if isinstance(v,float) else v) for k,v in (d.get('parity') or {}).items() if k!='vs'})
PY
