"""GPU timing helper: the per-rank shard of config 2 at 8 GPUs (32 trials = 640 segments) on ONE GPU, i.e. the
latency-bound regime of the strong-scaling run without the NCCL part.  Prints where an EM iteration's time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vlgp_b200 import core, gp
from vlgp_b200.core import Session
from vlgp_b200.gp import make_cholesky

ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 32
trials, params, config, c = bench.build_problem("config2")
trials = trials[:ntr]
make_cholesky(trials, params, config)
core.update_w(trials, params, config)
core.update_v(trials, params, config)
segs = bench.cut(trials, params, config)
make_cholesky(segs, params, config)
config["max_iter"] = config["min_iter"] = 1
quiet = open(os.devnull, "w")
sys.stdout = quiet
s = Session(segs, params)
for _ in range(3):
    core._em_iteration(s, segs, params, config)
sys.stdout = sys.__stdout__
ts, eng = s.ts, s.eng


def timeit(f, n=50):
    f()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    eng.sync()
    return (time.perf_counter() - t0) / n * 1e6


ts.hstep_prepare()
h = np.array([1.0, 0.01, 1e-4])
lat = list(range(5)); hh = np.tile(h, (5, 1))
print("segments %d" % len(segs))
print("hstep_objective_batch n=5: %8.1f us per round" % timeit(lambda: ts.hstep_objective_batch(lat, hh), 200))
print("hstep_objective n=1      : %8.1f us" % timeit(lambda: ts.hstep_objective(0, h), 200))
eng.profile_enable(0x4)
for _ in range(50):
    ts.hstep_objective_batch(lat, hh)
ms, n = eng.profile_get(2)
eng.profile_enable(0)
print("  segment kernel (events) : %8.1f us per launch" % (ms / n * 1e3))
print("hstep_prepare            : %8.1f us" % timeit(ts.hstep_prepare, 20))
print("make_cholesky (session)  : %8.1f us" % timeit(lambda: s.make_cholesky(params), 20))
print("norms                    : %8.1f us" % timeit(ts.norms, 50))

p2 = dict(params)
R = 10
config["hstep_rounds"] = []
t0 = time.perf_counter()
for _ in range(R):
    pp = dict(p2)
    gp._optimize_dev(s, pp, config)
tot = (time.perf_counter() - t0) / R
print("whole H-step (native optimiser): %.2f ms; %.1f device rounds -> %.1f us per round" % (
    tot * 1e3, np.mean(config["hstep_rounds"]), tot * 1e6 / max(np.mean(config["hstep_rounds"]), 1)))
print("mstep(25)                : %8.1f us" % timeit(lambda: ts.mstep(25), 10))
print("estep(25)                : %8.1f us" % timeit(lambda: ts.estep(25), 10))
for ov in (True, False, True, False):
    config["overlap_mh"] = ov
    sys.stdout = quiet
    t, sp = [], []
    for _ in range(12):
        eng.sync(); t0 = time.perf_counter()
        sp.append(core._em_iteration(s, segs, params, config)); ts.norms()
        t.append(time.perf_counter() - t0)
    sys.stdout = sys.__stdout__
    print("EM iteration overlap=%s: %.2f ms; split (E, M, H) = %s ms" % (
        ov, np.median(t) * 1e3, np.round(np.median(np.array(sp), axis=0) * 1e3, 2)))
s.close()
