#!/bin/bash
# A/B of the host-side shared-memory allreduce (bench.py, config 2) at N GPUs; prints value / split per arm.
N=${1:-2}
for arm in nccl shm; do
  if [ $arm = nccl ]; then export VLGP_NO_SHM=1; else unset VLGP_NO_SHM; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 > gpurun_out/shm_${arm}_$N.json 2> gpurun_out/shm_${arm}_$N.err
  python - gpurun_out/shm_${arm}_$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2f ms %.3f split %s e2e %.2f" % (d["value"], d["ms_per_step"], {k: round(v,2) for k,v in d["split_ms"].items() if k!="note"}, d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
