"""Full-trial inference (core.infer: E-step with Eniter := max_iter on the UNCUT trials, vlgp/core.py:260-266) on the
config-2 problem: device time of the launch sequence.  VLGP_NO_LONG_ESTEP=1 times the one-CTA-per-trial kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vlgp_b200 import core
from vlgp_b200.core import Session
from vlgp_b200.engine import get_engine
from vlgp_b200.gp import make_cholesky

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 20
trials, params, config, c = bench.build_problem(cfg)
make_cholesky(trials, params, config)
core.update_w(trials, params, config)
core.update_v(trials, params, config)
config["Eniter"] = n_iter
eng = get_engine()
with Session(trials, params) as s:
    s.ts.estep(2, config["dmu_bound"], config["method"])          # warm-up (allocations, module load)
    eng.profile_enable(0x1)
    for _ in range(3):
        s.refresh(trials, params)
        s.ts.estep(n_iter, config["dmu_bound"], config["method"])
    ms, n = eng.profile_get(0)
    eng.profile_enable(0)
    chk = s.ts.norms()
print("infer(%s, %d iterations): %.2f ms per call (%d calls)  norms %r  env %s" % (
    cfg, n_iter, ms / max(n, 1), n, [float(x) for x in np.ravel(chk)[:2]], {k: v for k, v in os.environ.items() if k.startswith("VLGP_")}))
