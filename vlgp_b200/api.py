"""Public entry points -- drop-in for vlgp/api.py: ``fit`` (:18-76), ``transform`` (:171-184), ``sample_posterior``
(:142-168).  Same arguments, same mutation of the caller's trial dicts, same ``{"trials", "params", "config"}`` return
surface (all host arrays float64)."""
from __future__ import annotations

import copy
import logging

import numpy as np

from .core import vem, update_w, update_v, infer, _echo
from .gp import make_cholesky
from .preprocess import get_params, get_config, fill_trials, fill_params, initialize
from .util import cut_trials

__all__ = ["fit", "sample_posterior", "transform"]

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

    _echo("Initializing")
    initialize(trials, params, config)
    _echo("Initialized")

    fill_params(params)
    fill_trials(trials)
    make_cholesky(trials, params, config)
    update_w(trials, params, config)
    update_v(trials, params, config)

    splits = cut_trials(trials, params, config)
    make_cholesky(splits, params, config)
    fill_trials(splits)

    params["initial"] = copy.deepcopy(params)

    _echo("Fitting")
    vem(splits, params, config)

    make_cholesky(trials, params, config)
    update_w(trials, params, config)
    update_v(trials, params, config)

    _echo("Inferring")
    infer(trials, params, config)
    _echo("Done")

    return {"trials": trials, "params": params, "config": config}


def transform(trials, params, config):
    """Infer the latent factors of new trials with a fitted model (trial lengths must already have a prior factor in
    ``params['cholesky']``, as in the reference)."""
    initialize(trials, params, config)
    fill_trials(trials)
    infer(trials, params, config)
    return trials


def sample_posterior(trial, params, nsamples, reg=1e-6):
    """Draw ``nsamples`` paths from the (full-covariance) posterior of one trial; returns (nsamples, bins, factors).
    Host-side like the reference (SURVEY.md section 8(f) item 2 lists the device version as future work)."""
    mu, w = trial["mu"], trial["w"]
    nbins, nfactors = mu.shape
    G = params["cholesky"][nbins]
    out = np.empty((nsamples, nbins, nfactors))
    eye = np.eye(nbins)
    for k in range(nfactors):
        K = G[k] @ G[k].T
        cov = np.linalg.inv(np.linalg.inv(K + reg * eye) + np.diag(w[:, k]))
        out[:, :, k] = np.random.multivariate_normal(mu[:, k], cov, size=nsamples)
    return out
