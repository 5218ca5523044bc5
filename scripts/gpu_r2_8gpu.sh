#!/bin/bash
mkdir -p gpurun_out
n=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
tail -3 gpurun_out/r2_bench_${n}gpu.err
python - <<'PY'
import json
for f in ('r2_bench_8gpu',):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'evals/step', d.get('hstep_evals_per_step'), 'e2e', d['e2e']['value'], d.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
