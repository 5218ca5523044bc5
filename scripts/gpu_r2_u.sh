#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mstep" 2>&1 | tail -3
python scripts/time_e2e.py > gpurun_out/r2u_time_e2e.txt 2>&1; cat gpurun_out/r2u_time_e2e.txt | tail -40
ncu --set full --clock-control none --import-source on -k regex:mstep_stats_tma_kernel -s 30 -c 1 -o gpurun_out/r2u_mstep_stats_tma_kernel -f python scripts/profile_driver.py 7 > gpurun_out/r2u_ncu_mstep_tma.log 2>&1
tail -3 gpurun_out/r2u_ncu_mstep_tma.log
