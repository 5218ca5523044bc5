#!/bin/bash
export VLGP_TIME_NTRIALS=32
python scripts/time_estep.py config2 4 8 2>&1 | tail -1 | cut -c1-100
VLGP_ESTEP_NO_FUSED=1 python scripts/time_estep.py config2 4 8 2>&1 | tail -1 | cut -c1-100
unset VLGP_TIME_NTRIALS
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_overlap.py -m gpu -x -q 2>&1 | tail -3
