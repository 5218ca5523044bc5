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


def sekernel(x, var, scale, jitter=1e-6):
    """Squared-exponential covariance ``var * exp(-(x_i - x_j)^2 / (2 scale^2)) + jitter * I`` (vlgp/gp.py:165-171; used by
    the GPFA branch only)."""
    x = np.asarray(x, dtype=float).reshape(-1, 1) / scale
    return var * np.exp(-0.5 * (x - x.T) ** 2) + np.eye(x.shape[0]) * jitter


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


def _lbfgsb(fun, x0, bounds, maxcor=10, ftol=2.2204460492503131e-09, gtol=1e-5, maxfun=15000, maxiter=15000, maxls=20):
    """scipy's L-BFGS-B with its defaults, driven through the same reverse-communication routine
    (``scipy.optimize._lbfgsb_py._lbfgsb.setulb``) as ``scipy.optimize.minimize(method="L-BFGS-B")`` but without the
    ``ScalarFunction`` / ``OptimizeResult`` machinery around every evaluation (about 0.1-0.8 ms of pure Python per
    evaluation, more than the device objective itself).  Same routine, same inputs, same iterates; falls back to the
    public ``minimize`` when the private entry point is missing or its signature has changed.
    Returns (x, f, nfev)."""
    x0 = np.asarray(x0, dtype=np.float64)
    bounds = np.asarray(bounds, dtype=np.float64)
    try:
        from scipy.optimize import _lbfgsb_py as _L

        setulb = _L._lbfgsb.setulb
        int_dtype = np.int64 if getattr(_L, "HAS_ILP64", False) else np.int32
        n = x0.size
        low, up = np.ascontiguousarray(bounds[:, 0]), np.ascontiguousarray(bounds[:, 1])
        nbd = np.full(n, 2, dtype=int_dtype)
        x = np.clip(x0, low, up).astype(np.float64)
        f = np.array(0.0, dtype=np.float64)
        g = np.zeros((n,), dtype=np.float64)
        m = maxcor
        wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        iwa = np.zeros(3 * n, dtype=int_dtype)
        task = np.zeros(2, dtype=int_dtype)
        ln_task = np.zeros(2, dtype=int_dtype)
        lsave = np.zeros(4, dtype=int_dtype)
        isave = np.zeros(44, dtype=int_dtype)
        dsave = np.zeros(29, dtype=np.float64)
        factr = ftol / np.finfo(float).eps
        nfev = nit = 0
        started = False
        while True:
            g = g.astype(np.float64)
            setulb(m, x, low, up, nbd, f, g, factr, gtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)
            started = True
            if task[0] == 3:
                f, g = fun(np.copy(x))
                nfev += 1
            elif task[0] == 1:
                nit += 1
                if nit >= maxiter:
                    task[0], task[1] = 5, 504
                elif nfev > maxfun:
                    task[0], task[1] = 5, 502
            else:
                break
        return x, float(f), nfev
    except (ImportError, AttributeError, TypeError):
        if "started" in locals() and locals().get("nfev", 0) > 0:
            raise
        from scipy.optimize import minimize

        res = minimize(fun, x0, jac=True, bounds=bounds, method="L-BFGS-B")
        return res.x, float(res.fun), int(res.nfev)


class _LbfgsbState:
    """Reverse-communication state of one L-BFGS-B run (the arrays scipy's driver keeps between setulb calls)."""

    def __init__(self, setulb, int_dtype, x0, bounds, maxcor=10, ftol=2.2204460492503131e-09, gtol=1e-5,
                 maxfun=15000, maxiter=15000, maxls=20):
        n = x0.size
        self.setulb, self.m, self.maxls, self.gtol = setulb, maxcor, maxls, gtol
        self.low, self.up = np.ascontiguousarray(bounds[:, 0]), np.ascontiguousarray(bounds[:, 1])
        self.nbd = np.full(n, 2, dtype=int_dtype)
        self.x = np.clip(x0, self.low, self.up).astype(np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros((n,), dtype=np.float64)
        m = maxcor
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=int_dtype)
        self.task = np.zeros(2, dtype=int_dtype)
        self.ln_task = np.zeros(2, dtype=int_dtype)
        self.lsave = np.zeros(4, dtype=int_dtype)
        self.isave = np.zeros(44, dtype=int_dtype)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.factr = ftol / np.finfo(float).eps
        self.nfev = self.nit = 0
        self.maxfun, self.maxiter = maxfun, maxiter
        self.done = False

    def advance(self):
        """Run the optimiser until it needs f, g at self.x (returns True) or terminates (returns False)."""
        while True:
            self.g = self.g.astype(np.float64)
            self.setulb(self.m, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.gtol, self.wa,
                        self.iwa, self.task, self.lsave, self.isave, self.dsave, self.maxls, self.ln_task)
            if self.task[0] == 3:
                return True
            if self.task[0] == 1:
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.maxfun:
                    self.task[0], self.task[1] = 5, 502
            else:
                self.done = True
                return False


def _lockstep_lbfgsb(ts, latents, initials, bounds, mask):
    """All per-latent L-BFGS-B runs advanced together in ONE host thread: every round collects the points the still
    active optimisers ask for and evaluates them in one batched device call.  Each optimiser sees exactly the values
    it would see alone (same setulb routine, same inputs), so the iterates equal the reference's sequential runs;
    the device sees one set of launches / one allreduce / one synchronisation per round.
    Returns [(x, f, nfev)] or None when scipy's private entry point is unavailable."""
    try:
        from scipy.optimize import _lbfgsb_py as _L

        setulb = _L._lbfgsb.setulb
        int_dtype = np.int64 if getattr(_L, "HAS_ILP64", False) else np.int32
        states = [_LbfgsbState(setulb, int_dtype, np.log(np.asarray(x0, dtype=np.float64)),
                               np.log(np.asarray(bounds, dtype=np.float64))) for x0 in initials]
        pending = [k for k, st in enumerate(states) if st.advance()]
    except (ImportError, AttributeError, TypeError):
        return None
    maskf = np.asarray(mask, dtype=float)
    while pending:
        hypers = np.array([np.exp(states[k].x) for k in pending])
        todo = list(range(len(pending)))
        ll = np.empty(len(pending))
        dll = np.empty(len(pending))
        while todo:      # the reference's retry when K is not PD (vlgp/gp.py:133-135)
            l_, d_, info = ts.hstep_objective_batch([latents[pending[i]] for i in todo], hypers[todo])
            again = []
            for q, i in enumerate(todo):
                if info[q] == 1:
                    hypers[i, 1] += np.log(10)
                    again.append(i)
                else:
                    ll[i], dll[i] = l_[q], d_[q]
            todo = again
        nxt = []
        for i, k in enumerate(pending):
            st = states[k]
            st.f = -ll[i]
            st.g = -(np.array([0.0, dll[i], 0.0]) * maskf)
            st.nfev += 1
            if st.advance():
                nxt.append(k)
        pending = nxt
    return [(st.x, float(st.f), st.nfev) for st in states]


def optimze1d(ts, latent, initial, bounds, mask, evaluate=None):
    """L-BFGS-B over log(sigma^2, omega, eps) of one latent (name kept from the reference, vlgp/gp.py:100-123).
    ``ts`` is a device TrialSet on which ``hstep_prepare`` has been called."""
    if evaluate is None:
        def evaluate(l, hyper):
            return ts.hstep_objective(l, hyper)
    x, fval, nfev = _lbfgsb(_objective(evaluate, latent, mask), np.log(initial), np.log(bounds))
    return np.exp(x), fval, nfev


def _optimize_scipy(ts, zdim, initials, bounds, mask):
    """The H-step driven by scipy's own L-BFGS-B from Python (VLGP_HSTEP_SCIPY=1; also what an engine stand-in without
    ``hstep_optimize`` gets): lockstep reverse communication, or threads / sequential ``minimize`` calls."""
    ts.hstep_prepare()
    results = [None] * zdim
    lock = None
    if not os.environ.get("VLGP_SEQUENTIAL_HSTEP") and not os.environ.get("VLGP_THREADED_HSTEP"):
        lock = _lockstep_lbfgsb(ts, list(range(zdim)), initials, bounds, mask)
    if lock is not None:
        results = [(np.exp(x), f, nf) for x, f, nf in lock]
    elif zdim > 1 and not os.environ.get("VLGP_SEQUENTIAL_HSTEP"):
        ev = _LockstepEvaluator(ts, zdim)
        errors = []

        def run(l):
            try:
                results[l] = optimze1d(ts, l, initials[l], bounds, mask, evaluate=ev.evaluate)
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
            results[l] = optimze1d(ts, l, initials[l], bounds, mask)
    return results


def _optimize_dev(s, params, config):
    from .util import blas_threads

    # scipy's L-BFGS-B routine solves 10 x 10 systems through LAPACK at every call: one BLAS thread (util.blas_threads)
    with blas_threads(0 if os.environ.get("VLGP_HSTEP_BLAS_THREADS") == "0" else 1):
        _optimize_dev_impl(s, params, config)


def _optimize_dev_impl(s, params, config):
    ts = s.ts
    zdim = params["zdim"]
    sigma = np.array(params["sigma"], dtype=float)
    omega = np.array(params["omega"], dtype=float)
    gp_noise = params["gp_noise"]
    lengths = getattr(ts, "lengths", None)
    if lengths is not None and len(set(np.asarray(lengths).tolist())) > 1:
        # the reference stacks the segments' mu / w (vlgp/gp.py:77-80): unequal lengths end in numpy's ValueError there
        raise ValueError("all input arrays must have the same shape")
    mask = np.array([0, 1, 0])
    bounds = ((1e-3, 1), config["omega_bound"], (gp_noise / 2, gp_noise * 2))
    initials = [(sigma[l] ** 2, omega[l], gp_noise) for l in range(zdim)]
    if hasattr(ts, "hstep_optimize") and not any(os.environ.get(k) for k in (
            "VLGP_HSTEP_SCIPY", "VLGP_SEQUENTIAL_HSTEP", "VLGP_THREADED_HSTEP")):
        # default: every round of every latent's L-BFGS-B inside one native call (csrc/hstep_opt.cu, csrc/lbfgsb.cuh)
        tol = config.get("hstep_collapse_tol", 1e-9)
        hyper, fval, nf, _, rounds = ts.hstep_optimize(list(range(zdim)), initials, bounds, mask, tol if tol else 0.0)
        results = [(hyper[l], float(fval[l]), int(nf[l])) for l in range(zdim)]
        config.setdefault("hstep_rounds", []).append(int(rounds))
    else:
        results = _optimize_scipy(ts, zdim, initials, bounds, mask)
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
