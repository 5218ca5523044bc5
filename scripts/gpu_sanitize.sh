#!/bin/bash
# compute-sanitizer over the kernels written in the last session (TMA-staged M-step statistics, wide-window H-step)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mstep_tma or wide_window_vs_oracle or mstep_golden" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
