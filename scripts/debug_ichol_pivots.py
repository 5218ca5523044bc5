"""GPU diagnostic: pivot sequences of the device ichol vs the NumPy oracle over many (n, omega)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import vlgp_oracle as orc
from vlgp_b200.engine import get_engine

eng = get_engine()
rng = np.random.default_rng(0)
L = 10
bad = 0
tot = 0
for n in (50, 137, 500, 733, 1000, 1337, 1999):
    omegas = np.exp(rng.uniform(np.log(5e-4), np.log(5e-2), L))
    params = dict(ydim=2, zdim=L, xdim=1, rank=50, gp_noise=1e-4, dt=1, likelihood=np.array(["poisson"] * 2),
                  sigma=np.ones(L), omega=omegas)
    eng.ensure_model(params)
    eng.push_params(params, which=("sigma", "omega"))
    with eng.new_trials([n]) as ts:
        ts.make_cholesky()
        G, piv, ncol = ts.get_cholesky(n, with_pivots=True)
    for l, om in enumerate(omegas):
        Gr, pr = orc.ichol_gauss(n, om, 50, return_pivots=True)
        tot += 1
        pd = piv[l][:ncol[l]]
        if len(pd) != len(pr) or not np.array_equal(pd, pr):
            bad += 1
            k = next((i for i in range(min(len(pd), len(pr))) if pd[i] != pr[i]), min(len(pd), len(pr)))
            # residuals of the oracle just before step k
            F = Gr[:, :k]
            resid = 1 - np.sum(F * F, axis=1)
            kk = np.exp(-om * (np.arange(n)[:, None] - np.arange(n)[None, :]) ** 2)
            err_d = np.abs(G[l] @ G[l].T - kk).max()
            err_r = np.abs(Gr @ Gr.T - kk).max()
            print("n=%d omega=%.6g first diff at step %d: dev %d ref %d  resid(dev)=%.17g resid(ref)=%.17g  ncol %d/%d  "
                  "|GG'-K| dev %.2e ref %.2e" % (n, om, k, pd[k] if k < len(pd) else -1, pr[k] if k < len(pr) else -1,
                                                  resid[pd[k]] if k < len(pd) else np.nan,
                                                  resid[pr[k]] if k < len(pr) else np.nan, len(pd), len(pr), err_d, err_r))
print("mismatching factorisations: %d of %d" % (bad, tot))
