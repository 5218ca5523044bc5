"""GPU timing helper: where the host-side time of one end-to-end vem() call goes (config 2)."""
import os, sys, time, cProfile, pstats
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
core.vem(segs, params, config)
sys.stdout = sys.__stdout__


def T(label, f, n=3):
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        r = f()
    print("%-28s %8.2f ms" % (label, (time.perf_counter() - t0) / n * 1e3))
    return r


s = T("Session()", lambda: Session(segs, params), 3)
ts = s.ts
ys = [t["y"] for t in segs]
T("  set_y_parts", lambda: ts.set_y_parts(ys))
mu = [t["mu"] for t in segs]
T("  set_state_parts(mu)", lambda: ts.set_state_parts(mu=mu))
T("  _bias_only", lambda: core._bias_only(segs, params))
T("  ensure+push params", lambda: (s.eng.ensure_model(params), s.eng.push_params(params)))
T("  set_cholesky", lambda: ts.set_cholesky(50, params["cholesky"][50]))
T("pull(all)", lambda: s.pull(segs))
T("  get_state_parts(mu)", lambda: ts.get_state_parts(mu=mu))
big = np.empty((ts.nbin, 5))
T("  get_state_parts(big)", lambda: ts.get_state_parts(w=[big]))
sys.stdout = open(os.devnull, "w")
t0 = time.perf_counter()
core.vem(segs, params, config)
dt = time.perf_counter() - t0
sys.stdout = sys.__stdout__
print("vem() whole call: %.2f ms" % (dt * 1e3))
cProfile.run("sys.stdout = open(os.devnull, 'w'); core.vem(segs, params, config); sys.stdout = sys.__stdout__", "/tmp/vem.prof")
pstats.Stats("/tmp/vem.prof").sort_stats("cumtime").print_stats(18)
