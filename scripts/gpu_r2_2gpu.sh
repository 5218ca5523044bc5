#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -30
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
tail -3 gpurun_out/r2_bench_${n}gpu.err
VLGP_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_${n}gpu_nccl.json 2> /dev/null
done
python - <<'PY'
import json
for f in ('r2_bench_2gpu','r2_bench_2gpu_nccl'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], d['split_ms'], 'evals/step', d.get('hstep_evals_per_step'), 'e2e', d['e2e']['value'], d.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
