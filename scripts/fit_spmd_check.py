"""SPMD fit() check, launched with torch.distributed.run on >= 2 ranks: every rank calls fit() on the same trials; all
ranks must return the same posterior for every trial, equal to the reference's golden runs:
  * Hstep=False (tests/golden/fit_fixed_omega.npz): 1e-7;
  * defaults, H-step on (tests/golden/fit_tutorial.npz): the pivot-flip scale of the single-process test (5e-4, omega 1e-5).
With VLGP_COMM_NO_NCCL=1 VLGP_DEVICE=0 the ranks share ONE GPU and talk through peer memory + shared memory only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vlgp_b200 as vlgp
from vlgp_b200 import dist
from vlgp_b200.synth import make_trials

eng = dist.init_from_env()


def run(golden, tol, tol_omega, **kw):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", golden + ".npz")))
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    sys.stdout = open(os.devnull, "w")
    res = vlgp.fit(trials, 3, max_iter=3, min_iter=3, **kw)
    sys.stdout = sys.__stdout__
    err = {}
    for k in ("mu", "v"):
        got = np.stack([t[k] for t in res["trials"]])
        err[k] = float(np.max(np.abs(got - g[k])) / np.max(np.abs(g[k])))
    for k in ("a", "b"):
        err[k] = float(np.max(np.abs(res["params"][k] - g[k])) / np.max(np.abs(g[k])))
    err_om = float(np.max(np.abs(res["params"]["omega"] - g["omega"])) / np.max(np.abs(g["omega"]))) if "omega" in g else 0.0
    s = float(np.sum(np.stack([t["mu"] for t in res["trials"]]))) + float(np.sum(res["params"]["a"]))
    mx = eng.allreduce(np.array([s, -s]), op="max")             # identical on every rank
    same = abs(mx[0] + mx[1]) <= 1e-12 * max(1.0, abs(mx[0]))
    worst = max(err.values())
    print("rank %d world %d %s worst rel err %.3e omega %.1e same_on_all_ranks %s peer_memory %s"
          % (dist.rank(), dist.world_size(), golden, worst, err_om, same, eng.peer_memory), flush=True)
    return worst < tol and err_om < tol_omega and same


ok = run("fit_fixed_omega", 1e-7, 1e-12, Hstep=False)
ok = run("fit_tutorial", 5e-4, 1e-5) and ok
dist.barrier()
sys.exit(0 if ok else 1)
