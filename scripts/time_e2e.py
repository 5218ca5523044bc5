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


def mk():
    x = Session(segs, params)
    x.close()


T("Session() + close", mk, 5)
lengths = [tr["y"].shape[0] for tr in segs]


def nt():
    t = core.get_engine().new_trials(lengths)
    t.free()


T("  new_trials + free", nt, 5)
T("  _shared_rows", lambda: core._shared_rows(segs), 5)
T("  lengths list", lambda: [tr["y"].shape[0] for tr in segs], 5)
s = Session(segs, params)
ts = s.ts
ys = [t["y"] for t in segs]
T("  set_y_parts", lambda: ts.set_y_parts(ys))
mu = [t["mu"] for t in segs]
T("  set_state_parts(mu)", lambda: ts.set_state_parts(mu=mu))
vv = [t["v"] for t in segs]
ww = [t["w"] for t in segs]
T("  set_state_parts(mu,v,w)", lambda: ts.set_state_parts(mu=mu, v=vv, w=ww))
T("  list comprehension x3", lambda: ([t["mu"] for t in segs], [t["v"] for t in segs], [t["w"] for t in segs]))
T("  _bias_only", lambda: core._bias_only(segs, params))
T("  ensure+push params", lambda: (s.eng.ensure_model(params), s.eng.push_params(params)))
T("  set_cholesky", lambda: ts.set_cholesky(50, params["cholesky"][50]))
T("pull(all)", lambda: s.pull(segs))
T("  norms", lambda: ts.norms())
T("  get_state_parts(mu)", lambda: ts.get_state_parts(mu=mu))
big = np.empty((ts.nbin, 5))
T("  get_state_parts(big)", lambda: ts.get_state_parts(w=[big]))
sys.stdout = open(os.devnull, "w")
t0 = time.perf_counter()
core.vem(segs, params, config)
dt = time.perf_counter() - t0
sys.stdout = sys.__stdout__
for rep in range(3):
    sys.stdout = open(os.devnull, "w")
    t0 = time.perf_counter()
    core.vem(segs, params, config)
    dt = time.perf_counter() - t0
    sys.stdout = sys.__stdout__
    print("vem() again: %.2f ms" % (dt * 1e3), {k: [round(x * 1e3, 2) for x in v] for k, v in config["runtime"].items() if k != "it"})
print("vem() whole call: %.2f ms" % (dt * 1e3), {k: [round(x * 1e3, 2) for x in v] for k, v in config["runtime"].items() if k != "it"})
with Session(segs, params) as s2:
    sys.stdout = open(os.devnull, "w")
    for i in range(3):
        s2.eng.sync(); t0 = time.perf_counter()
        sp = core._em_iteration(s2, segs, params, config)
        s2.eng.sync(); dt = time.perf_counter() - t0
        sys.stdout = sys.__stdout__
        print("resident iteration %d: %.2f ms, split %s" % (i, dt * 1e3, [round(x * 1e3, 2) for x in sp]))
        sys.stdout = open(os.devnull, "w")
    sys.stdout = sys.__stdout__
cProfile.run("sys.stdout = open(os.devnull, 'w'); core.vem(segs, params, config); sys.stdout = sys.__stdout__", "/tmp/vem.prof")
pstats.Stats("/tmp/vem.prof").sort_stats("cumtime").print_stats(18)
