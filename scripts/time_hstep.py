"""GPU timing helper: latency of single / batched H-step objective evaluations and of the other per-iteration calls
on the config-2 problem."""
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
for _ in range(3):
    core._em_iteration(s, segs, params, config)
sys.stdout = sys.__stdout__
ts, eng = s.ts, s.eng
ts.hstep_prepare()
h = np.array([1.0, 0.01, 1e-4])


def timeit(f, n=50):
    f()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    eng.sync()
    return (time.perf_counter() - t0) / n * 1e6


print("hstep_objective n=1      : %8.1f us" % timeit(lambda: ts.hstep_objective(0, h)))
for n in (2, 5):
    lat = list(range(n))
    hh = np.tile(h, (n, 1))
    print("hstep_objective_batch n=%d: %8.1f us" % (n, timeit(lambda: ts.hstep_objective_batch(lat, hh))))
eng.profile_enable(0x4)
for _ in range(20):
    ts.hstep_objective(0, h)
ms, n = eng.profile_get(2)
print("segment kernel alone (events): %.1f us per launch" % (ms / n * 1e3))
eng.profile_enable(0x4)
lat = list(range(5)); hh = np.tile(h, (5, 1))
for _ in range(20):
    ts.hstep_objective_batch(lat, hh)
ms, n = eng.profile_get(2)
print("segment kernel batch of 5 (events): %.1f us per launch" % (ms / n * 1e3))
eng.profile_enable(0)
print("hstep_prepare            : %8.1f us" % timeit(ts.hstep_prepare, 10))
print("norms                    : %8.1f us" % timeit(ts.norms, 20))
print("make_cholesky (session)  : %8.1f us" % timeit(lambda: s.make_cholesky(params), 10))
t0 = time.perf_counter()
sys.stdout = open(os.devnull, "w")
core._hstep_dev(s, segs, params, config)
sys.stdout = sys.__stdout__
print("whole H-step: %.1f ms, nfev %s" % ((time.perf_counter() - t0) * 1e3, config["hstep_nfev"][-1]))
os.environ["VLGP_SEQUENTIAL_HSTEP"] = "1"
t0 = time.perf_counter()
core._hstep_dev(s, segs, params, config)
print("whole H-step sequential: %.1f ms, nfev %s" % ((time.perf_counter() - t0) * 1e3, config["hstep_nfev"][-1]))
eng.profile_enable(0x2)
ts.mstep(25)
ms, n = eng.profile_get(1)
print("mstep stats kernel (events): %.1f us per launch" % (ms / n * 1e3))
eng.profile_enable(0)
print("mstep(25)                : %8.1f us" % timeit(lambda: ts.mstep(25), 5))
print("estep(25)                : %8.1f us" % timeit(lambda: ts.estep(25), 3))
s.close()
