#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float32" 2>&1 | grep -E "passed|failed|^E  +assert|err" | head
python - <<'PY'
import sys, copy, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from test_gpu_parity import _problem, _cfg
from conftest import relerr
from vlgp_b200 import core
from oracle import vlgp_oracle as orc
for (N, L, nt, sc) in [(200, 10, 2, 0.3), (200, 10, 2, 0.1), (200, 10, 2, 0.2)]:
    segs, params = _problem(11, nt, 1000, N, L, a_scale=sc)
    s_ref = copy.deepcopy(segs); orc.estep(s_ref, copy.deepcopy(params), _cfg(Eniter=25))
    s64 = copy.deepcopy(segs); core.estep(s64, copy.deepcopy(params), _cfg(Eniter=25))
    s32 = copy.deepcopy(segs); core.estep(s32, copy.deepcopy(params), _cfg(Eniter=25, dtype="float32"))
    e = [relerr(a["mu"], b["mu"]) for a, b in zip(s64, s_ref)]
    e32 = [relerr(a["mu"], b["mu"]) for a, b in zip(s32, s_ref)]
    clipped = [float(np.mean(np.abs(b["dmu"]) >= 5.0 - 1e-9)) for b in s_ref]
    print(N, L, sc, "f64 max err %.2e" % max(e), "f32 max err %.2e median %.2e" % (max(e32), np.median(e32)), "n bad", sum(x > 1e-8 for x in e), "of", len(e),
          "max |dmu| last it (oracle) %.2e" % max(np.abs(b["dmu"]).max() for b in s_ref))
PY
