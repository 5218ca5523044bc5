#!/bin/bash
# compute-sanitizer over the bordered H-step kernel (all its template sizes) and the whole H-step test group
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hstep" > gpurun_out/sanitize2_memcheck.log 2>&1
echo "== memcheck"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize2_memcheck.log | sort | uniq -c | head -6
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lengths_vs_oracle and (17 or 18 or 26 or 33 or 42 or 49 or 50)" > gpurun_out/sanitize2_racecheck.log 2>&1
echo "== racecheck"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize2_racecheck.log | sort | uniq -c | head -6
