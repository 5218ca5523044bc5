#!/bin/bash
# 8-GPU session: config 2 at the driver's settings with 2 / 3 / 4 M-step iterations pumped per H-step round, config 4
mkdir -p gpurun_out
n=8
run() {  # tag config steps warmup port [env]
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $5 bench.py --gpus $n --config $2 --steps $3 --warmup $4 --no-cpu > gpurun_out/r2x_bench_${n}gpu_$1.json 2> gpurun_out/r2x_bench_${n}gpu_$1.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2x_bench_${n}gpu_$1.json'))
    print('$1', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'E', round(d['roofline']['ms_per_launch'],3), 'e2e', round(d['e2e']['value'],1), {k: (float('%.2g' % v) if isinstance(v,float) else v) for k,v in (d.get('parity') or {}).items() if k!='vs'})
except Exception as e:
    print('$1 ERR', e)
PY
}
VLGP_MSTEP_PUMP=2 run config2_pump2 config2 20 5 29611
VLGP_MSTEP_PUMP=3 run config2_pump3 config2 20 5 29612
VLGP_MSTEP_PUMP=4 run config2_pump4 config2 20 5 29613
run config4 config4 10 5 29614
