#!/bin/bash
for n in 2 3 4 6 8; do
python scripts/time_estep.py config2 $n 4 2>&1 | tail -1 | cut -c1-60
VLGP_ESTEP_NO_FUSED=1 python scripts/time_estep.py config2 $n 4 2>&1 | tail -1 | cut -c1-80
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
