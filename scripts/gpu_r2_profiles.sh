#!/bin/bash
# Round-2 profiles at HEAD: bench line, launch list of 3 EM iterations, ncu --set full of the three top kernels.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -c 600 gpurun_out/r2m_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2m_launches.csv python scripts/profile_driver.py 3 > gpurun_out/r2m_launches.log 2>&1
for k in estep_seg_kernel hstep_segment_dmma_kernel mstep_stats_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r2m_$k -f python scripts/profile_driver.py 6 > gpurun_out/r2m_ncu_$k.log 2>&1
  tail -2 gpurun_out/r2m_ncu_$k.log
done
ls -la gpurun_out | grep r2m
