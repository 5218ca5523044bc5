#!/bin/bash
for kb in 8192 2048 1024; do
echo "== stage $kb KB"
VLGP_STAGE_KB=$kb python scripts/time_e2e.py 2>&1 | grep -E "^Session|^pull|set_y_parts|set_state_parts\(mu,v,w\)|vem\(\) again" | tr '\n' ';' | cut -c1-420
echo
done
for v in msu1 msu4; do
VLGP_B200_LIB=$PWD/vlgp_b200/variants/libvlgp_b200_$v.so python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$v', round(d['value'],2), 'M-step ms', round(d['roofline_mstep']['ms_per_launch'],4), round(d['roofline_mstep']['frac'],3))"
done
python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('default', round(d['value'],2), 'M-step ms', round(d['roofline_mstep']['ms_per_launch'],4), round(d['roofline_mstep']['frac'],3))"
