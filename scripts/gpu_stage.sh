#!/bin/bash
nproc
for cfg in "8 8192" "16 8192" "32 8192" "8 2048" "16 2048" "16 1024"; do
set -- $cfg
echo "== threads $1 ystage $2 KB"
VLGP_HOST_THREADS=$1 VLGP_YSTAGE_KB=$2 python scripts/time_e2e.py 2>&1 | grep -E "^Session|^pull|set_y_parts|set_state_parts\(mu,v,w\)|vem\(\) again" | tr '\n' ';' | cut -c1-330
echo
done
