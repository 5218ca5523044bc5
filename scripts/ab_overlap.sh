#!/bin/bash
# A/B of the overlapped M-/H-step (bench.py, config 2) at 1 and N GPUs on one box; prints value / split per arm.
N=${1:-2}
show() { python - "$1" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2f ms %.3f split %s e2e %.2f" % (d["value"], d["ms_per_step"], {k: round(v,2) for k,v in d["split_ms"].items() if k!="note"}, d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for arm in seq ovl; do
  if [ $arm = seq ]; then export VLGP_NO_OVERLAP=1; else unset VLGP_NO_OVERLAP; fi
  timeout 200 python bench.py --no-cpu --steps 10 > gpurun_out/ab_${arm}_1.json 2> gpurun_out/ab_${arm}_1.err; show gpurun_out/ab_${arm}_1.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 > gpurun_out/ab_${arm}_$N.json 2> gpurun_out/ab_${arm}_$N.err; show gpurun_out/ab_${arm}_$N.json
done
