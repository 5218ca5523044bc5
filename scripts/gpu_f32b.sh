#!/bin/bash
for n in 1 5; do
python scripts/time_estep.py config2 $n 6 2>&1 | tail -1 | cut -c1-120
VLGP_TIME_DTYPE=float32 python scripts/time_estep.py config2 $n 6 2>&1 | tail -1 | cut -c1-120
done
VLGP_ESTEP_NO_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float32" 2>&1 | grep -E "passed|failed|assert 0" | head
