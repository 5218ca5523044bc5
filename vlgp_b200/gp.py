"""GP prior factors and hyperparameter optimisation -- drop-in for vlgp/gp.py.

``make_cholesky`` (:150-162) runs the batched pivoted incomplete Cholesky on the GPU; ``optimize`` (:65-97) keeps
scipy's L-BFGS-B on the host (host code stays Python) but every objective/gradient evaluation
(``construct_posterior_cov`` + ``elbo``, :12-62,126-147) is one batched device call over all segments.
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine

__all__ = ["make_cholesky", "optimize", "optimze1d"]


def make_cholesky(trials, params, config=None):
    """params['cholesky'] = {length: (zdim, length, rank)} for every unique trial length (REPLACES the dict)."""
    eng = get_engine()
    eng.ensure_model(params)
    eng.push_params(params, which=("sigma", "omega"))
    lengths = sorted({int(tr["y"].shape[0]) for tr in trials})
    with eng.new_trials(lengths) as ts:          # one placeholder trial per unique length: only the factors are used
        ts.make_cholesky()
        params["cholesky"] = {t: ts.get_cholesky(t) for t in lengths}


def _objective(ts, latent, mask):
    mask = np.asarray(mask, dtype=float)

    def fun(x):
        hyper = np.exp(x)
        while True:
            ll, dll, info = ts.hstep_objective(latent, hyper)
            if info != 1:
                break
            hyper[1] += np.log(10)          # the reference's retry when K is not PD (vlgp/gp.py:133-135)
        grad = np.array([0.0, dll, 0.0]) * mask
        return -ll, -grad

    return fun


def optimze1d(ts, latent, initial, bounds, mask):
    """L-BFGS-B over log(sigma^2, omega, eps) of one latent (name kept from the reference, vlgp/gp.py:100-123).
    ``ts`` is a device TrialSet on which ``hstep_prepare`` has been called."""
    from scipy.optimize import minimize

    res = minimize(_objective(ts, latent, mask), np.log(initial), jac=True, bounds=np.log(bounds))
    return np.exp(res.x), res.fun, res.nfev


def _optimize_dev(s, params, config):
    ts = s.ts
    zdim = params["zdim"]
    sigma = np.array(params["sigma"], dtype=float)
    omega = np.array(params["omega"], dtype=float)
    gp_noise = params["gp_noise"]
    ts.hstep_prepare()
    nfev = []
    for l in range(zdim):
        initial = (sigma[l] ** 2, omega[l], gp_noise)
        bounds = ((1e-3, 1), config["omega_bound"], (gp_noise / 2, gp_noise * 2))
        (sigmasq, omega_new, _), _, nf = optimze1d(ts, l, initial, bounds, mask=np.array([0, 1, 0]))
        if not np.any(np.isclose(omega_new, config["omega_bound"])):
            omega[l] = omega_new
        sigma[l] = np.sqrt(sigmasq)
        nfev.append(nf)
    params["sigma"] = sigma
    params["omega"] = omega
    config.setdefault("hstep_nfev", []).append(nfev)
    s.make_cholesky(params)


def optimize(trials, params, config):
    """Optimise the GP hyperparameters of every latent on the given (equal-length) segments."""
    from .core import Session

    with Session(trials, params, upload_factors=False) as s:
        _optimize_dev(s, params, config)
