"""E-step launch time on the config-2 problem after a few EM iterations (steady-state omega); one process per setting of
the environment switches the kernel reads.  usage: time_estep.py [config] [em_iterations] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vlgp_b200 import core
from vlgp_b200.core import Session
from vlgp_b200.engine import get_engine
from vlgp_b200.gp import make_cholesky

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
n_em = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_launch = int(sys.argv[3]) if len(sys.argv) > 3 else 6
trials, params, config, c = bench.build_problem(cfg)
if os.environ.get("VLGP_TIME_NTRIALS"):
    trials = trials[:int(os.environ["VLGP_TIME_NTRIALS"])]
make_cholesky(trials, params, config)
core.update_w(trials, params, config)
core.update_v(trials, params, config)
segs = bench.cut(trials, params, config)
make_cholesky(segs, params, config)
config["max_iter"] = config["min_iter"] = 1
if os.environ.get("VLGP_TIME_DTYPE"):
    config["dtype"] = os.environ["VLGP_TIME_DTYPE"]
out = sys.stdout
sys.stdout = open(os.devnull, "w")
eng = get_engine()
with Session(segs, params) as s:
    for _ in range(n_em):
        core._em_iteration(s, segs, params, config)
    s.push_params(params) if hasattr(s, "push_params") else None
    eng.profile_enable(0x1)
    for _ in range(n_launch):
        eng.flush_l2()
        core.estep(segs, params, config, session=s)
    ms, n = eng.profile_get(0)
    eng.profile_enable(0)
    chk = s.ts.norms()
sys.stdout = out
print("estep %.3f ms per launch (%d launches)  env: %s  norms %r" % (ms / max(n, 1), n, {k: v for k, v in os.environ.items() if k.startswith("VLGP_")}, [float(x) for x in np.ravel(chk)[:3]]))
