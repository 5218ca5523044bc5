"""Host-side set-up of the three dicts the reference passes around (trials / params / config).

Mirrors the behaviour of vlgp/preprocess.py (get_config :84-112, get_params :49-81, initialize :4-46, fill_trials
:115-120, fill_params :123-125) including its use of the *global* NumPy RNG (one ``np.random.choice`` per initialize),
so that a seeded run draws the same numbers as the reference.  Nothing here is on the hot path: it stays NumPy/sklearn
on the host, exactly like the reference (SURVEY.md section 8(f) item 1 lists a device initialiser as future work).
"""
from __future__ import annotations

import numpy as np

__all__ = ["get_config", "get_params", "initialize", "fill_trials", "fill_params"]

_CONFIG_DEFAULTS = (
    ("constrain_loading", "fro"),
    ("constrain_latent", False),
    ("use_hessian", True),
    ("eps", 1e-8),
    ("tol", 1e-8),
    ("min_iter", 5),
    ("method", "VB"),
    ("learning_rate", 1.0),
    ("max_iter", 20),
    ("Eniter", 25),
    ("Mniter", 25),
    ("Hstep", True),
    ("da_bound", 5.0),
    ("db_bound", 5.0),
    ("dmu_bound", 5.0),
    ("omega_bound", (5e-4, 5e-2)),
    ("window", 50),
    ("saving_interval", 60 * 30),
    ("callbacks", None),
    ("parallel", False),
    ("dtype", "float64"),          # not in the reference: "float32" = single-precision E-step rate passes (engine.set_precision)
)


def get_config(**kwargs):
    """Default configuration overridden by the recognised keyword arguments; unknown ones are dropped silently."""
    config = {k: ([] if k == "callbacks" else v) for k, v in _CONFIG_DEFAULTS}
    for k in list(config):
        if k in kwargs:
            config[k] = kwargs[k]
    return config


def get_params(trials, zdim, **kwargs):
    """Initial parameter dict.  ``kwargs`` must carry ``omega_bound`` (fit passes config's)."""
    ydim = trials[0]["y"].shape[-1]
    lik = kwargs.get("lik", "poisson")
    lik = np.asarray(lik if isinstance(lik, list) else [lik] * ydim)
    return {
        "ydim": ydim,
        "zdim": zdim,
        "xdim": max(kwargs.get("history", 0), 1),
        "a": kwargs.get("a", None),
        "b": kwargs.get("b", None),
        "noise": kwargs.get("noise", np.ones(ydim)),
        "sigma": kwargs.get("sigma", np.ones(zdim)),
        "omega": kwargs.get("omega", np.full(zdim, kwargs["omega_bound"][1], dtype=float)),
        "rank": 50,
        "gp_noise": 1e-4,
        "dt": 1,
        "likelihood": lik,
    }


def _blas_thread_cap():
    """Cap of the BLAS thread pool while the factor analysis runs.

    FactorAnalysis.fit is a handful of randomized SVDs of a tall, skinny matrix (bins/10 x neurons): every product is a
    few Mflop, and OpenBLAS with one thread per core spends its time waking and spinning threads -- measured on the
    256 x 1000 x 100 workload: 2.0 s with 8 threads, 0.19 s with 4, 0.32 s with 1 (1.49 of fit()'s 1.75 s wall time on
    the GPU box went here).  The arithmetic is sklearn's own either way; the summation order inside GEMM depends on the
    thread count, as it does for the reference between two machines (loading agrees to 1e-14, latents to 1e-12).
    VLGP_INIT_BLAS_THREADS overrides the cap of 4; 0 leaves the pool alone."""
    import os

    from .util import blas_threads

    try:
        cap = int(os.environ.get("VLGP_INIT_BLAS_THREADS", "4"))
    except ValueError:
        cap = 4
    return blas_threads(max(cap, 0))


def initialize(trials, params, config, defer_mu=False):
    """Factor-analysis initialisation of loading / bias / noise and of every trial's posterior mean
    (vlgp/preprocess.py:4-46).  ``defer_mu=True`` (used by fit): when the factor model is fitted here and no trial
    brings its own ``mu``, the per-trial ``transform`` calls are left out and the projection ``(mean, P, C)`` with
    ``mu = ((y - mean) @ P) @ C`` is returned instead, for the engine to evaluate on the device where y lives anyway
    (TrialSet.project_y); returns None when every ``mu`` has been set on the host."""
    with _blas_thread_cap():
        return _initialize(trials, params, config, defer_mu)


def _rows(ys, lengths, pick):
    """Rows ``pick`` of the concatenation of the blocks ``ys`` without building the concatenation."""
    if len(ys) == 1:
        return ys[0][pick, :]
    starts = np.concatenate([[0], np.cumsum(lengths)])
    owner = np.searchsorted(starts, pick, side="right") - 1
    out = np.empty((pick.size, ys[0].shape[1]), dtype=np.result_type(*[y.dtype for y in ys]))
    order = np.argsort(owner, kind="stable")
    bounds = np.searchsorted(owner[order], np.arange(len(ys) + 1))
    for i in np.flatnonzero(np.diff(bounds)):
        sel = order[bounds[i]:bounds[i + 1]]
        out[sel] = ys[i][pick[sel] - starts[i], :]
    return out


def _initialize(trials, params, config, defer_mu=False):
    from sklearn.decomposition import FactorAnalysis

    zdim, xdim = params["zdim"], params["xdim"]
    ys = [tr["y"] for tr in trials]
    lengths = [y.shape[0] for y in ys]
    nbin, ydim = int(sum(lengths)), ys[0].shape[-1]
    pick = np.random.choice(nbin, max(nbin // 10, 50))      # with replacement, like the reference

    projection = None
    if params.get("transform") is None:
        y_pick = _rows(ys, lengths, pick)                   # y_all[pick, :] of the reference, without the 200 MB concat
        fa = FactorAnalysis(n_components=zdim, random_state=0)
        z = fa.fit_transform(y_pick)
        params["transform"] = fa.transform
        if params.get("a") is None:
            params["a"] = fa.components_
        if params.get("b") is None:
            mean = np.concatenate(ys, axis=0).mean(axis=0, keepdims=True) if len(ys) == 1 or not _same_dtype(ys) else \
                _mean_rows(ys, nbin)
            params["b"] = np.log(np.maximum(mean, config["eps"]))
        if params.get("noise") is None:
            params["noise"] = np.var(y_pick - z @ fa.components_, ddof=0, axis=0)
        if defer_mu and params["a"] is fa.components_ and all(tr.get("mu") is None for tr in trials):
            # FactorAnalysis.transform: ((X - mean) @ Wpsi') @ cov_z with Wpsi = components / noise_variance
            wpsi = fa.components_ / fa.noise_variance_
            cov_z = np.linalg.inv(np.eye(zdim) + wpsi @ fa.components_.T)
            projection = (np.array(fa.mean_, dtype=float), np.ascontiguousarray(wpsi.T), cov_z)

    to_latent = params["transform"]
    for tr in trials:
        nt = tr["y"].shape[0]
        if tr.get("mu") is None and projection is None:
            tr["mu"] = to_latent(tr["y"])
        if tr.get("x") is None:
            tr["x"] = np.ones((nt, xdim, ydim))
            _mark_ones(tr["x"])
        tr["w"] = np.zeros((nt, zdim))
        tr["v"] = np.zeros((nt, zdim))
    return projection


def _same_dtype(ys):
    return len({y.dtype for y in ys}) == 1


def _mean_rows(ys, nbin):
    """Column means of the concatenation of the blocks (one pass per block, no concatenation)."""
    tot = np.zeros(ys[0].shape[1])
    for y in ys:
        tot += y.sum(axis=0)
    return (tot / nbin)[None, :]


def _mark_ones(x):
    """Tell the engine's regressor check (core._all_ones) that this freshly made array is the all-ones bias column, so
    that it is not scanned again (205 MB at 256 trials x 1000 bins x 100 neurons)."""
    try:
        from .core import _remember_ones

        _remember_ones(x)
    except Exception:  # pragma: no cover - bookkeeping only
        pass


def fill_trials(trials):
    """``cut`` index and zero ``w`` / ``v`` / ``dmu`` where missing (vlgp/preprocess.py:115-120).  The missing arrays of
    one key are handed out as row blocks of ONE zero array when every ``mu`` is a 2-D float64 array of the same width
    (thousands of segments: one allocation instead of one ``zeros_like`` each)."""
    for i, tr in enumerate(trials):
        tr["cut"] = i
    for key in ("w", "v", "dmu"):
        need = [tr for tr in trials if key not in tr]
        if not need:
            continue
        mus = [tr["mu"] for tr in need]
        width = {(m.shape[1] if isinstance(m, np.ndarray) and m.ndim == 2 and m.dtype == np.float64 else None)
                 for m in mus}
        if len(need) > 1 and len(width) == 1 and None not in width:
            rows = [m.shape[0] for m in mus]
            block = np.zeros((sum(rows), width.pop()))
            r0 = 0
            for tr, n in zip(need, rows):
                tr[key] = block[r0:r0 + n]
                r0 += n
        else:
            for tr, m in zip(need, mus):
                tr[key] = np.zeros_like(m)


def fill_params(params):
    if "da" not in params:
        params["da"] = np.zeros_like(params["a"])
    if "db" not in params:
        params["db"] = np.zeros_like(params["b"])
