#!/bin/bash
for st in 0 4000 8000 12000 18000 24000 36000; do
  VLGP_ESTEP_STAGGER=$st python scripts/time_estep.py config2 4 6 2>&1 | tail -1
done
