"""GPFA branch: the Gaussian-likelihood EM of the reference (vlgp/gpfa.py) behind the same functions.

``fit`` / ``prepare`` / ``em`` / ``infer`` / ``leastsq`` / ``loglik`` / ``make_prior`` keep the reference's names, argument
meaning and return values (vlgp/gpfa.py:11-158).  The E-step, which the reference evaluates as an ``(n * ydim)``-square dense
solve with ``np.kron`` matrices, runs on the device through the push-through form described in csrc/gpfa.cu; so does the
accumulation of the least-squares normal equations of the M-step.  The host keeps what is tiny: the ``(zdim * n)``-square
matrix P of the iteration and the ``(zdim + 1)``-square solve.

Two properties of the reference are reproduced on purpose, because a drop-in has to return its numbers:
  * ``bigR = np.kron(np.eye(n), R)`` is built ONCE, before the loop (vlgp/gpfa.py:31): the E-step of every iteration uses
    the R that ``em`` was called with, never the M-step's;
  * that ``bigR`` is ordered (time, neuron) while ``bigC`` and the residual are ordered (neuron, time)
    (vlgp/gpfa.py:30-31,41-42): the noise applied to (neuron j, bin t) is ``R[(j * n + t) % ydim]``.
Both are invisible with the ``R = I`` that ``prepare`` hands to ``em``.
"""
import time

import numpy as np
from numpy import linalg

from .gp import sekernel
from .preprocess import get_config, get_params, initialize, fill_params, fill_trials
from .util import cut_trials

__all__ = ["make_prior", "em", "infer", "leastsq", "loglik", "fit", "prepare"]


def _echo(msg):
    print(msg)


def make_prior(trials, n_factors, dt, var, scale):
    """Squared-exponential prior covariance of every trial, ``trial['K']`` (vlgp/gpfa.py:11-17)."""
    for trial in trials:
        n = trial["y"].shape[0]
        trial["K"] = sekernel(np.arange(n) * dt, var, scale)


def _noise_table(Rdiag, n, ydim):
    """rho[t, j] = 1 / (noise the reference applies to neuron j at bin t), see the module docstring."""
    idx = (np.arange(ydim)[None, :] * n + np.arange(n)[:, None]) % ydim
    return 1.0 / np.asarray(Rdiag, dtype=float)[idx]


def _posterior_operator(C, rho, K):
    """P = (I + bigK S)^-1 bigK with S = bigC' bigR^-1 bigC (block-diagonal in time); vector index ``l * n + t``."""
    zdim, n = C.shape[0], K.shape[0]
    Q = np.einsum("lj,kj,tj->lkt", C, C, rho)
    S = np.zeros((zdim * n, zdim * n))
    t = np.arange(n)
    for l in range(zdim):
        for k in range(zdim):
            S[l * n + t, k * n + t] = Q[l, k]
    bigK = np.kron(np.eye(zdim), K)
    return linalg.solve(np.eye(zdim * n) + bigK @ S, bigK)


def _model(ydim, zdim, dt=1):
    return {"ydim": ydim, "zdim": zdim, "xdim": 1, "rank": 1, "gp_noise": 1e-4, "dt": dt,
            "likelihood": np.array(["gaussian"] * ydim)}


def em(y, C, d, R, K, max_iter):
    """EM of the Gaussian GPFA model ``p(y|z) = N(zC + d, R)``, ``p(z) = N(0, K)`` on stacked equal-length segments
    ``y`` (trial, time, dim) -- vlgp/gpfa.py:20-56.  Returns ``(z, C, d, R)``."""
    from .engine import get_engine

    y = np.asarray(y)
    m, n, ydim = y.shape
    C = np.array(C, dtype=float)
    zdim = C.shape[0]
    d = np.asarray(d, dtype=float).reshape(1, ydim)
    rho = _noise_table(np.diag(R), n, ydim)                 # from the R passed in, for every iteration
    eng = get_engine()
    eng.ensure_model(_model(ydim, zdim))
    z = np.zeros((m, n, zdim))
    with eng.new_trials([n] * m) as ts:
        ts.set_y(np.ascontiguousarray(y.reshape(m * n, ydim), dtype=np.float64))
        for i in range(max_iter):
            t0 = time.perf_counter()
            ts.gpfa_estep(C, d, rho, _posterior_operator(C, rho, K))
            s, _, count = ts.latent_moments()
            ts.latent_affine(shift=s / count)               # z -= z.mean(axis=(0, 1))
            t1 = time.perf_counter()
            ztz, zty, yy = ts.gpfa_stats()
            coef = linalg.solve(ztz, zty)                   # lstsq(Z1, Y) through its normal equations
            r = yy - 2.0 * np.sum(coef * zty, axis=0) + np.einsum("aj,ab,bj->j", coef, ztz, coef)
            C, d = coef[:-1, :], coef[[-1], :]
            R = np.diag(r ** 2)
            C = C / linalg.norm(C)
            t2 = time.perf_counter()
            _echo("Iteration {:4d}, E-step {:.2f}s, M-step {:.2f}s".format(i + 1, t1 - t0, t2 - t1))
        if max_iter > 0:
            z = ts.get_state(("mu",))["mu"].reshape(m, n, zdim)
    return z, C, d, R


def infer(trials, C, d, R):
    """Posterior mean of every trial under fixed (C, d, R), ``trial['mu']`` (vlgp/gpfa.py:59-76); each trial uses its own
    ``trial['K']``.  The operator P is (zdim * length)-square: meant for the lengths the reference can handle itself."""
    from .engine import get_engine

    C = np.asarray(C, dtype=float)
    zdim, ydim = C.shape
    d = np.asarray(d, dtype=float).reshape(1, ydim)
    eng = get_engine()
    eng.ensure_model(_model(ydim, zdim))
    for i, trial in enumerate(trials):
        t0 = time.perf_counter()
        n = trial["y"].shape[0]
        rho = _noise_table(np.diag(R), n, ydim)
        with eng.new_trials([n]) as ts:
            ts.set_y(np.ascontiguousarray(trial["y"], dtype=np.float64))
            ts.gpfa_estep(C, d, rho, _posterior_operator(C, rho, trial["K"]))
            trial["mu"] = ts.get_state(("mu",))["mu"].reshape(n, zdim)
        _echo("Trial {:d}, {:.2f}s".format(i, time.perf_counter() - t0))


def leastsq(Y, Z, constant=True):
    """``Y = Z C + d`` by least squares; returns (C, d, residual sums of squares) -- vlgp/gpfa.py:79-85 (host; the EM loop
    accumulates the same normal equations on the device)."""
    if constant:
        Z = np.column_stack([Z, np.ones(Z.shape[0])])
    C, r, *_ = linalg.lstsq(Z, Y, rcond=None)
    return C[:-1, :], C[[-1], :], r


def loglik(y, z, C, d, R, var, scale, dt):
    """The reference's (unnormalised, sign-flipped) log-likelihood expression, vlgp/gpfa.py:88-101 (host)."""
    zdim, ydim = C.shape
    m, n, _ = y.shape
    K = sekernel(np.arange(n) * dt, var, scale)
    bigK = np.kron(np.eye(zdim), K)
    r = y - z @ C - d[None, :]
    r = r @ (1 / np.sqrt(R))
    Z = z.transpose((0, 2, 1)).reshape(m, -1, 1)
    return np.sum(r ** 2) + np.sum(Z.transpose((0, 2, 1)) @ linalg.solve(bigK[None, ...], Z)) + m * linalg.slogdet(bigK)[1]


def fit(trials, n_factors, **kwargs):
    """GPFA fit: returns ``(y, z, C, d, R)`` like vlgp/gpfa.py:104-124."""
    y, C, d, R, K = prepare(trials, n_factors, **kwargs)
    _echo("Fitting")
    z, C, d, R = em(y, C, d, R, K, kwargs["max_iter"])
    return y, z, C, d, R


def prepare(trials, n_factors, **kwargs):
    """Initialisation, prior and stacked segments for ``em`` (vlgp/gpfa.py:127-158): ``(y, C, d, R, K)``."""
    config = get_config(**kwargs)
    kwargs["omega_bound"] = config["omega_bound"]
    params = get_params(trials, n_factors, **kwargs)
    _echo("Initializing")
    t0 = time.perf_counter()
    initialize(trials, params, config)
    _echo("Initialized {:.2f}s".format(time.perf_counter() - t0))
    fill_params(params)
    params["R"] = np.eye(trials[0]["y"].shape[1])
    dt, var, scale = kwargs["dt"], kwargs["var"], kwargs["scale"]
    fill_trials(trials)
    make_prior(trials, n_factors=n_factors, dt=dt, var=var, scale=scale)
    segments = cut_trials(trials, params, config)
    y = np.stack([segment["y"] for segment in segments])
    C, d, R = params["a"], params["b"], params["R"]
    n = config["window"]
    K = sekernel(np.arange(n) * dt, var, scale)
    return y, C, d, R, K
