#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_trial or config3_shape or transform or reference_api or fit_tutorial" 2>&1 | tail -5
python scripts/time_infer.py config2 20 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k3_launches.csv python scripts/time_infer.py config2 4 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/k3_launches.csv | head -8
