"""GP prior factors and hyperparameter optimisation -- drop-in for vlgp/gp.py.

``make_cholesky`` (:150-162) runs the batched pivoted incomplete Cholesky on the GPU; ``optimize`` (:65-97) keeps
scipy's L-BFGS-B on the host (host code stays Python) but every objective/gradient evaluation
(``construct_posterior_cov`` + ``elbo``, :12-62,126-147) is one batched device call over all segments.
"""
from __future__ import annotations

import os
import threading

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


class _LockstepEvaluator:
    """Batches the objective evaluations of the per-latent L-BFGS-B runs.

    The reference optimises the latents one after the other (vlgp/gp.py:82-92); the runs are independent, so here each
    runs in its own host thread and blocks in ``evaluate`` until every still-active run has asked for its next point;
    the last one to arrive issues ONE batched device call for all of them.  Each optimiser sees exactly the values it
    would see alone, so the iterates are those of the sequential runs; the device sees L times fewer launches,
    allreduces and synchronisations.  Batch composition depends only on the optimisers' own progress, so every rank
    of a multi-GPU run issues the same sequence of collectives."""

    def __init__(self, ts, n_active):
        self.ts = ts
        self.cond = threading.Condition()
        self.pending = {}
        self.results = {}
        self.active = n_active
        self.error = None

    def _flush(self):
        lats = sorted(self.pending)
        try:
            ll, dll, info = self.ts.hstep_objective_batch(lats, np.array([self.pending[l] for l in lats]))
            for k, l in enumerate(lats):
                self.results[l] = (float(ll[k]), float(dll[k]), int(info[k]))
        except BaseException as e:  # noqa: BLE001 - re-raised in every waiting thread
            self.error = e
        self.pending.clear()
        self.cond.notify_all()

    def evaluate(self, latent, hyper):
        with self.cond:
            if self.error is not None:
                raise self.error
            self.pending[latent] = np.array(hyper, dtype=float)
            if len(self.pending) >= self.active:
                self._flush()
            while latent not in self.results and self.error is None:
                self.cond.wait()
            if self.error is not None:
                raise self.error
            return self.results.pop(latent)

    def retire(self):
        with self.cond:
            self.active -= 1
            if self.pending and len(self.pending) >= self.active:
                self._flush()


def _objective(evaluate, latent, mask):
    mask = np.asarray(mask, dtype=float)

    def fun(x):
        hyper = np.exp(x)
        while True:
            ll, dll, info = evaluate(latent, hyper)
            if info != 1:
                break
            hyper[1] += np.log(10)          # the reference's retry when K is not PD (vlgp/gp.py:133-135)
        grad = np.array([0.0, dll, 0.0]) * mask
        return -ll, -grad

    return fun


def optimze1d(ts, latent, initial, bounds, mask, evaluate=None):
    """L-BFGS-B over log(sigma^2, omega, eps) of one latent (name kept from the reference, vlgp/gp.py:100-123).
    ``ts`` is a device TrialSet on which ``hstep_prepare`` has been called."""
    from scipy.optimize import minimize

    if evaluate is None:
        def evaluate(l, hyper):
            return ts.hstep_objective(l, hyper)
    res = minimize(_objective(evaluate, latent, mask), np.log(initial), jac=True, bounds=np.log(bounds))
    return np.exp(res.x), res.fun, res.nfev


def _optimize_dev(s, params, config):
    ts = s.ts
    zdim = params["zdim"]
    sigma = np.array(params["sigma"], dtype=float)
    omega = np.array(params["omega"], dtype=float)
    gp_noise = params["gp_noise"]
    ts.hstep_prepare()
    mask = np.array([0, 1, 0])
    bounds = ((1e-3, 1), config["omega_bound"], (gp_noise / 2, gp_noise * 2))
    results = [None] * zdim
    if zdim > 1 and not os.environ.get("VLGP_SEQUENTIAL_HSTEP"):
        ev = _LockstepEvaluator(ts, zdim)
        errors = []

        def run(l):
            try:
                results[l] = optimze1d(ts, l, (sigma[l] ** 2, omega[l], gp_noise), bounds, mask, evaluate=ev.evaluate)
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
            finally:
                ev.retire()

        threads = [threading.Thread(target=run, args=(l,), daemon=True) for l in range(zdim)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
    else:
        for l in range(zdim):
            results[l] = optimze1d(ts, l, (sigma[l] ** 2, omega[l], gp_noise), bounds, mask)
    nfev = []
    for l in range(zdim):
        (sigmasq, omega_new, _), _, nf = results[l]
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
