#!/bin/bash
python scripts/time_e2e.py 2>&1 | head -50
