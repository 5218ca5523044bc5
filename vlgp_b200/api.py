"""Public entry points -- drop-in for vlgp/api.py: ``fit`` (:18-76), ``transform`` (:171-184), ``sample_posterior``
(:142-168).  Same arguments, same mutation of the caller's trial dicts, same ``{"trials", "params", "config"}`` return
surface (all host arrays float64)."""
from __future__ import annotations

import copy
import logging

import numpy as np

from . import dist
from .core import Session, vem, update_w, update_v, infer, _echo
from .gp import make_cholesky
from .preprocess import get_params, get_config, fill_trials, fill_params, initialize
from .util import cut_trials

__all__ = ["fit", "sample_posterior", "posterior_cov", "transform", "map2vi", "fastfit", "resume"]

logger = logging.getLogger(__name__)


def fit(trials, n_factors, **kwargs):
    """Fit a vLGP model.

    :param trials: list of dicts with at least ``y`` of shape (bins, channels); optional ``x``, ``mu``
    :param n_factors: number of latent factors
    :param kwargs: ``lik``, ``history``, ``a``, ``b``, ``noise``, ``sigma``, ``omega`` and any config key
        (vlgp/preprocess.py:59-74,85-106); unknown keys are ignored
    :return: ``{"trials": trials, "params": params, "config": config}``
    """
    config = get_config(**kwargs)
    logger.info("\n".join("{} : {}".format(k, v) for k, v in config.items()))

    kwargs["omega_bound"] = config["omega_bound"]
    params = get_params(trials, n_factors, **kwargs)
    # limits of the device kernels, checked before any work is done (the reference has none of them)
    if config["Hstep"] and config["window"] and int(config["window"]) > 160 and config["max_iter"] > 0:
        raise ValueError("window=%d with Hstep=True: the H-step kernels hold a window of at most 160 bins "
                         "(pass Hstep=False or a window <= 160)" % int(config["window"]))
    if int(params["zdim"]) > 12:
        raise ValueError("n_factors=%d: the device kernels are instantiated for at most 12 latents" % int(params["zdim"]))

    _echo("Initializing")
    # the factor model is fitted on the host (scikit-learn, like the reference); the per-trial transform that gives the
    # initial posterior means runs on the device, where y is uploaded anyway (see below)
    projection = initialize(trials, params, config, defer_mu=True)
    _echo("Initialized")

    fill_params(params)
    if projection is None:
        fill_trials(trials)

    world, rank = dist.world_size(), dist.rank()
    if world > len(trials):
        raise ValueError("%d ranks for %d trials: every rank needs at least one trial" % (world, len(trials)))
    if world > 1:
        # every rank ran the RNG-consuming host set-up above on its own; make rank 0's outcome everybody's, whatever
        # the state of each process's global NumPy generator was: loading, bias, noise (in place: params["a"] may BE
        # the factor model's components_), the projection / initial means, and the generator state the window cuts are drawn from
        from .util import assign_inplace

        for key in ("a", "b", "noise"):
            assign_inplace(params, key, dist.broadcast_from_root(params[key]))
        if projection is not None:
            projection = tuple(dist.broadcast_from_root(p) for p in projection)
        else:
            for tr in trials:
                tr["mu"][...] = dist.broadcast_from_root(tr["mu"])
        kind, keys, pos, has_gauss, cached = np.random.get_state()
        st = dist.broadcast_from_root(np.concatenate([keys.astype(float), [float(pos), float(has_gauss), cached]]))
        np.random.set_state((kind, st[:-3].astype(np.uint32), int(st[-3]), int(st[-2]), float(st[-1])))

    # Multi-GPU (one process per GPU, vlgp_b200.dist.init_from_env() called first): SPMD -- every rank makes the same
    # call on the same trials, the host set-up above is replicated, each rank runs the device work of its contiguous
    # shard of trials (the M-/H-step statistics are allreduced inside vem), and at the end every rank holds the
    # posterior of every trial.  With one process this is exactly the reference's sequence (vlgp/api.py:49-71).
    lo, hi = dist.shard_bounds(len(trials), world, rank)
    mine = trials[lo:hi] if world > 1 else trials

    # the uncut trials stay on the device from here to the final infer (one upload of y instead of five)
    if projection is not None:
        L = params["zdim"]
        for tr in trials:                    # placeholders: the device fills in this rank's, the final gather the others'
            tr["mu"] = np.zeros((tr["y"].shape[0], L))
    full = Session(mine, params, upload_factors=False)
    try:
        if projection is not None:
            full.ts.project_y(*projection)
            full.pull(mine, ("mu",))
            fill_trials(trials)
        return _fit_on_device(trials, mine, lo, hi, world, full, params, config)
    finally:
        full.close()


def _fit_on_device(trials, mine, lo, hi, world, full, params, config):
    full.make_cholesky(params)                       # make_cholesky(trials): params["cholesky"][T] per trial length
    update_w(mine, params, config, session=full)
    update_v(mine, params, config, session=full)

    splits = cut_trials(trials, params, config)      # all trials: the RNG stream equals the single-process run's
    if world > 1 and config["window"]:
        import math

        per_trial = [math.ceil(tr["y"].shape[0] / config["window"]) for tr in trials]
        first = sum(per_trial[:lo])
        splits = splits[first:first + sum(per_trial[lo:hi])]
    elif world > 1:
        splits = mine
    make_cholesky(splits, params, config)
    fill_trials(splits)

    params["initial"] = copy.deepcopy(params)

    _echo("Fitting")
    vem(splits, params, config)

    # the segments are views of the trials: vem wrote mu and v through them; parameters and omega changed
    full.refresh(mine, params, which=("mu", "v"))
    full.make_cholesky(params)
    update_w(mine, params, config, session=full)
    update_v(mine, params, config, session=full)

    _echo("Inferring")
    infer(mine, params, config, session=full)
    if world > 1:
        _gather_trials(trials, lo, hi, params)
        make_cholesky(trials, params, config)        # params["cholesky"] for every trial length, on every rank
    _echo("Done")

    return {"trials": trials, "params": params, "config": config}


def _gather_trials(trials, lo, hi, params):
    """Give every rank the (mu, v, w, dmu) of every trial: zero-filled concatenation + one sum-allreduce per key."""
    from .engine import get_engine

    eng = get_engine()
    L = params["zdim"]
    lengths = [tr["y"].shape[0] for tr in trials]
    starts = np.concatenate([[0], np.cumsum(lengths)])
    for key in ("mu", "v", "w", "dmu"):
        buf = np.zeros((int(starts[-1]), L))
        for i in range(lo, hi):
            buf[starts[i]:starts[i + 1]] = trials[i][key]
        eng.allreduce_bulk(buf)
        for i, tr in enumerate(trials):
            if lo <= i < hi:
                continue
            val = buf[starts[i]:starts[i + 1]]
            if key in ("mu", "v") and isinstance(tr.get(key), np.ndarray) and tr[key].shape == val.shape:
                tr[key][...] = val
            else:
                tr[key] = val.copy()


def transform(trials, params, config):
    """Infer the latent factors of new trials with a fitted model (trial lengths must already have a prior factor in
    ``params['cholesky']``, as in the reference)."""
    initialize(trials, params, config)
    fill_trials(trials)
    infer(trials, params, config)
    return trials


def posterior_cov(trial, params, reg=1e-6):
    """Full posterior covariance of every latent of one trial, (factors, bins, bins):
    ``inv(inv(K + reg I) + diag(w))`` with ``K = G G'`` (vlgp/api.py:160-166; ``reg=0`` gives util.posterior_cov,
    vlgp/util.py:541-547).  Computed on the device from the rank-r prior factor (csrc/postcov.cu)."""
    from .engine import get_engine

    mu, w = np.asarray(trial["mu"], dtype=float), np.asarray(trial["w"], dtype=float)
    nbins, nfactors = mu.shape
    G = params["cholesky"][nbins]
    eng = get_engine()
    model = {"ydim": 1, "zdim": nfactors, "xdim": 1, "rank": G.shape[2], "gp_noise": params.get("gp_noise", 1e-4),
             "dt": params.get("dt", 1), "likelihood": np.array(["poisson"])}
    eng.ensure_model(model)
    out = np.empty((nfactors, nbins, nbins))
    with eng.new_trials([nbins]) as ts:
        ts.set_cholesky(nbins, G)
        ts.set_state(mu=mu, v=np.zeros_like(mu), w=w)
        for k in range(nfactors):
            out[k] = ts.posterior_cov(0, k, reg)
    return out


def sample_posterior(trial, params, nsamples, reg=1e-6):
    """Draw ``nsamples`` paths from the (full-covariance) posterior of one trial; returns (nsamples, bins, factors)
    (vlgp/api.py:142-168).  The bins x bins covariance of every latent comes from the device (``posterior_cov``); the
    draws use ``np.random.multivariate_normal`` on the host like the reference, so they consume the global generator
    the same way.  (The reference's own draws are not reproducible beyond that: the covariance has ~bins - rank
    singular values at the level of ``reg``, and the SVD inside multivariate_normal picks an arbitrary basis of that
    subspace -- a 1e-10 change of the matrix changes individual samples at O(1) while their distribution is the same.)"""
    mu = np.asarray(trial["mu"], dtype=float)
    nbins, nfactors = mu.shape
    cov = posterior_cov(trial, params, reg)
    out = np.empty((nsamples, nbins, nfactors))
    for k in range(nfactors):
        out[:, :, k] = np.random.multivariate_normal(mu[:, k], cov[k], size=nsamples)
    return out


def map2vi(trials, C, d, **kwargs):
    """vLGP inference started from a GPFA solution: loading ``C``, bias ``log(d)``, then five E-step iterations on the
    uncut trials (vlgp/api.py:79-105).  Like the reference it updates ``trials`` in place and returns what ``resume``
    returns, i.e. None."""
    n_factors = trials[0]["mu"].shape[-1]
    config = get_config(**kwargs)
    logger.info("\n".join(["{} : {}".format(k, v) for k, v in config.items()]))
    config["callbacks"] = config["callbacks"] or []
    kwargs["omega_bound"] = config["omega_bound"]
    params = get_params(trials, n_factors, **kwargs)
    params["a"] = C
    params["b"] = np.log(d)
    make_cholesky(trials, params, config)
    update_w(trials, params, config)
    update_v(trials, params, config)
    config["max_iter"] = 5
    return resume(trials, params, config)


def fastfit(trials, n_factors, dt, var, scale, max_iter=20, **kwargs):
    """GPFA (MAP) fit followed by vLGP inference from it (vlgp/api.py:108-119)."""
    from . import gpfa

    omega = np.full(n_factors, 0.5 / ((scale / dt) ** 2))
    y, C, d, R, K = gpfa.prepare(trials, n_factors, dt=dt, var=var, scale=scale)
    z, C, d, R = gpfa.em(y, C, d, R, K, max_iter)
    return map2vi(trials, C, d, omega=omega, **kwargs)


def resume(trials, params, config):
    """Full-trial inference under the given parameters (vlgp/api.py:122-125)."""
    _echo("Inferring")
    infer(trials, params, config)
    _echo("Done")
