#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_overlap.py -m gpu -x -q 2>&1 | tail -3
python scripts/time_small_shard.py 32 2>&1 | grep -E "estep\(25\)" | head -3
python scripts/time_estep.py config2 5 4 2>&1 | tail -1 | cut -c1-60
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --n-trials 32 > gpurun_out/r2_small_1gpu.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2q_bench.json 2>/dev/null
python - <<'PY'
import json
for f in ('r2_small_1gpu','r2q_bench'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, round(d['value'],1), round(d['ms_per_step'],3), 'E kernel', round(d['roofline']['ms_per_launch'],3), {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'})
PY
