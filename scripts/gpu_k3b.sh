#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k3_launches.csv python scripts/time_infer.py config2 4 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/k3_launches.csv
