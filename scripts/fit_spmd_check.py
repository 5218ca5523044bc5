"""SPMD fit() check, launched with torch.distributed.run on >= 2 GPUs: every rank calls fit() on the same trials;
all ranks must return the same posterior for every trial, equal to the reference's golden run (Hstep=False: 1e-7)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vlgp_b200 as vlgp
from vlgp_b200 import dist
from vlgp_b200.synth import make_trials

eng = dist.init_from_env()
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                              "fit_fixed_omega.npz")))
trials = make_trials(10, 200, 30, 3, seed=0)
np.random.seed(0)
sys.stdout = open(os.devnull, "w")
res = vlgp.fit(trials, 3, max_iter=3, min_iter=3, Hstep=False)
sys.stdout = sys.__stdout__
err = {}
for k in ("mu", "v", "w"):
    got = np.stack([t[k] for t in res["trials"]])
    err[k] = float(np.max(np.abs(got - g[k])) / np.max(np.abs(g[k])))
for k in ("a", "b"):
    err[k] = float(np.max(np.abs(res["params"][k] - g[k])) / np.max(np.abs(g[k])))
worst = max(err.values())
# identical on every rank
chk = np.array([float(np.sum(np.stack([t["mu"] for t in res["trials"]]))), -float(np.sum(np.stack([t["mu"] for t in res["trials"]])))])
mx = eng.allreduce(chk.copy(), op="max")
same = abs(mx[0] + mx[1]) < 1e-9 * max(1.0, abs(mx[0]))
print("rank %d world %d worst rel err %.3e same_on_all_ranks %s" % (dist.rank(), dist.world_size(), worst, same), flush=True)
dist.barrier()
sys.exit(0 if (worst < 1e-7 and same) else 1)
