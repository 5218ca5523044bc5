#!/bin/bash
# 8-GPU runs of the BASELINE configurations that name 8 GPUs (config 2 strong scaling, configs 4 and 5).
N=${1:-8}
for cfg in config2 config4 config5; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --config $cfg > gpurun_out/scale${N}_$cfg.json 2> gpurun_out/scale${N}_$cfg.err
  echo "$cfg rc=$?"; python - gpurun_out/scale${N}_$cfg.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.2f ms %.3f split %s e2e %.2f" % (d["value"], d["ms_per_step"], {k: round(v,2) for k,v in d["split_ms"].items() if k!="note"}, d["e2e"]["value"]))
except Exception as e:
    print("  FAILED", e)
PY
done
