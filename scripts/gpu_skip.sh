#!/bin/bash
for sk in 0 1 2 4 8 12 14 15; do
VLGP_DEBUG_SKIP=$sk python scripts/time_estep.py config2 5 4 2>&1 | tail -1 | cut -c1-110
VLGP_DEBUG_SKIP=$sk VLGP_ESTEP_NO_FUSED=1 python scripts/time_estep.py config2 5 4 2>&1 | tail -1 | cut -c1-110
done
