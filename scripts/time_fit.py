"""GPU timing helper: whole vlgp_b200.fit() on BASELINE config 2 (host initialisation included), with the split the
reference itself reports (config["runtime"]) and a cProfile of the call."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vlgp_b200 as vlgp
from vlgp_b200.synth import make_trials

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 5
trials = make_trials(256, 1000, 100, 5, seed=0)
np.random.seed(0)
sys.stdout = open(os.devnull, "w")
vlgp.fit(make_trials(4, 200, 100, 5, seed=1), 5, max_iter=1, min_iter=1)      # warm-up (library load, pools)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
res = vlgp.fit(trials, 5, max_iter=n_iter, min_iter=n_iter)
pr.disable()
dt = time.perf_counter() - t0
sys.stdout = sys.__stdout__
rt = res["config"]["runtime"]
print("fit(256 x 1000 x 100 x 5, max_iter=%d): %.2f s wall; vem iterations: %s ms" % (
    n_iter, dt, ["%.1f" % (1e3 * x) for x in rt["em_elapsed"]]))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(22)
print("\n".join(s.getvalue().splitlines()[4:34]))
