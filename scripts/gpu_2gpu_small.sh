#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --n-trials 32 > gpurun_out/r2_small_1gpu.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --n-trials 64 > gpurun_out/r2_small_2gpu.json 2> gpurun_out/r2_small_2gpu.err
VLGP_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --n-trials 64 > gpurun_out/r2_small_2gpu_nop2p.json 2> /dev/null
python - <<'PY'
import json
for f in ('r2_small_1gpu','r2_small_2gpu','r2_small_2gpu_nop2p'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'evals/step', d.get('hstep_evals_per_step'), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
