#!/bin/bash
# usage: gpu_r2_scale.sh N   -- bench.py at the driver's settings on N GPUs of this box
n=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu > gpurun_out/r2z_bench_${n}gpu.json 2> gpurun_out/r2z_bench_${n}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2z_bench_${n}gpu.json'))
print('N=$n', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'E kernel', round(d['roofline']['ms_per_launch'],3), 'e2e', round(d['e2e']['value'],1), {k: (float('%.2g' % v) if isinstance(v,float) else v) for k,v in (d.get('parity') or {}).items() if k!='vs'})
PY
