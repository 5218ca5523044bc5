#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for n in 1 5; do
python scripts/time_estep.py config2 $n 6 2>&1 | tail -1
VLGP_ESTEP_NO_FUSED=1 python scripts/time_estep.py config2 $n 6 2>&1 | tail -1
done
VLGP_DEBUG_SKIP=14 python scripts/time_estep.py config2 5 4 2>&1 | tail -1 | cut -c1-110
