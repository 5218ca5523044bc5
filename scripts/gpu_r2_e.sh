#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:estep_seg_kernel -s 5 -c 1 -o gpurun_out/r2_estep_fast -f python scripts/profile_driver.py 6 > gpurun_out/r2_ncu_estep_fast.log 2>&1
tail -2 gpurun_out/r2_ncu_estep_fast.log
