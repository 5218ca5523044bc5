#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_overlap.py -m gpu -x -q 2>&1 | grep -vE "^Iteration|^Trial|Initializ|Fitting|Inferring|Done" | tail -4
for n in 1 5; do
python scripts/time_estep.py config2 $n 6 2>&1 | tail -1 | cut -c1-120
VLGP_ESTEP_NO_SPLIT_ROWS=1 python scripts/time_estep.py config2 $n 6 2>&1 | tail -1 | cut -c1-120
done
