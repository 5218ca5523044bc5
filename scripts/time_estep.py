"""GPU timing helper: E-step (25 iterations) on the config-2 problem after a few EM iterations (steady-state omega)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vlgp_b200 import core
from vlgp_b200.core import Session
from vlgp_b200.gp import make_cholesky

trials, params, config, c = bench.build_problem("config2")
make_cholesky(trials, params, config)
core.update_w(trials, params, config)
core.update_v(trials, params, config)
segs = bench.cut(trials, params, config)
make_cholesky(segs, params, config)
config["max_iter"] = config["min_iter"] = 1
sys.stdout = open(os.devnull, "w")
s = Session(segs, params)
for _ in range(4):
    core._em_iteration(s, segs, params, config)
sys.stdout = sys.__stdout__
ts, eng = s.ts, s.eng
print("ncol", [int((np.abs(params["cholesky"][50][l]).sum(axis=0) > 0).sum()) for l in range(5)])
st = ts.get_state()
for skip in (0, 1, 2, 4, 8, 14, 15):
    os.environ["VLGP_DEBUG_SKIP"] = str(skip)
    ts.set_state(st["mu"], st["v"], st["w"])
    ts.estep(25)
    ts.set_state(st["mu"], st["v"], st["w"])
    eng.sync(); t0 = time.perf_counter()
    ts.estep(25)
    eng.sync(); print("skip=%2d estep(25): %.2f ms" % (skip, (time.perf_counter() - t0) * 1e3))
os.environ["VLGP_DEBUG_SKIP"] = "0"
ts.set_state(st["mu"], st["v"], st["w"])
eng.sync(); t0 = time.perf_counter()
for _ in range(3):
    ts.mstep(25)
eng.sync(); print("mstep(25): %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
s.close()
