#!/bin/bash
# First GPU call of round 2: whole GPU suite at HEAD, A/B of the E-step rate-pass variants, launch list and ncu --set full
# captures of the three top kernels (E-step segments, H-step DMMA segments, M-step statistics).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
cat gpurun_out/r2_pytest.log
bash scripts/ab_estep_variants.sh bench 2>&1 | tail -12
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_head.json 2> gpurun_out/r2_bench_head.err
tail -c 3000 gpurun_out/r2_bench_head.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python scripts/profile_driver.py 3 > gpurun_out/r2_launches.log 2>&1
for k in estep_seg_kernel hstep_segment_dmma_kernel mstep_stats_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r2_$k -f python scripts/profile_driver.py 2 > gpurun_out/r2_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -20
