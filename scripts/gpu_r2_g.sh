#!/bin/bash
mkdir -p gpurun_out
for name in scalar ilp1_flat ilp1_warp ilp2_flat default; do
  lib=$PWD/vlgp_b200/variants/libvlgp_b200_$name.so
  VLGP_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$name', round(d['value'],2), 'EM-iter/s; E-step', round(d['roofline']['ms_per_launch'],3), 'ms per launch; split', {k: round(v,2) for k,v in d['split_ms'].items() if k!='note'})"
done
