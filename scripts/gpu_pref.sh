#!/bin/bash
python scripts/time_e2e.py 2>&1 | grep -E "vem\(\)"
VLGP_NO_PREFETCH=1 python scripts/time_e2e.py 2>&1 | grep -E "vem\(\)"
