#!/bin/bash
mkdir -p gpurun_out
n=8
run() {  # config steps tag
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $n --config $1 --steps $2 --warmup 3 --no-cpu > gpurun_out/r2p_bench_${n}gpu_$3.json 2> gpurun_out/r2p_bench_${n}gpu_$3.err
tail -2 gpurun_out/r2p_bench_${n}gpu_$3.err | cut -c1-300
}
run config2 20 config2 29611
run config4 10 config4 29612
run config5 6 config5 29613
python - <<'PY'
import json
for f in ('config2','config4','config5'):
    try:
        d=json.load(open('gpurun_out/r2p_bench_8gpu_%s.json'%f))
        print(f, round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'}, 'evals/step', round(d.get('hstep_evals_per_step',0),1), 'e2e', round(d['e2e']['value'],1), {k: (round(v,2) if isinstance(v,float) and v>1e-3 else v) for k,v in (d.get('parity') or {}).items() if k!='vs'})
    except Exception as e: print(f, 'ERR', e)
PY
