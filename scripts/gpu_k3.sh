#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_trial or config3_shape or transform or reference_api or fit_tutorial" 2>&1 | tail -12
python scripts/time_infer.py config2 20 2>&1 | tail -1
VLGP_NO_LONG_ESTEP=1 python scripts/time_infer.py config2 20 2>&1 | tail -1
