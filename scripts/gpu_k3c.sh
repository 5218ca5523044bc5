#!/bin/bash
VLGP_K3_DEBUG=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k3_launches_dbg.csv python scripts/time_infer.py config2 4 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/k3_launches_dbg.csv 2>/dev/null | head -5
