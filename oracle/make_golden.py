"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference, via oracle/ref_shim.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree is not on the GPU box):

    python oracle/make_golden.py

Every fixture stores the exact inputs next to the reference's outputs so that the tests never need the reference.
Covers SURVEY.md section 8(c) items (i)-(vii).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, vlgp_oracle as orc  # noqa: E402
from vlgp_b200.synth import make_trials  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_problem(ref, seed, n_trials, T, N, L, lik=None, window=None, **cfgkw):
    """Reference-format (trials, params, config) with random but reproducible state (no FactorAnalysis)."""
    rng = np.random.default_rng(seed)
    trials = make_trials(n_trials, T, N, L, seed=seed + 1000)
    config = ref.preprocess.get_config(**cfgkw)
    kwargs = {"omega_bound": config["omega_bound"]}
    if lik is not None:
        kwargs["lik"] = lik
    params = ref.preprocess.get_params(trials, L, **kwargs)
    params["a"] = 0.3 * rng.standard_normal((L, N))
    params["b"] = np.log(np.maximum(np.mean(np.concatenate([t["y"] for t in trials]), axis=0, keepdims=True), 1e-2))
    params["noise"] = 0.5 + rng.random(N)
    params["omega"] = np.exp(rng.uniform(np.log(2e-3), np.log(4e-2), L))
    for t in trials:
        n = t["y"].shape[0]
        if lik is not None:
            gauss = np.asarray(lik) == "gaussian"
            t["y"][:, gauss] = t["y"][:, gauss] + 0.3 * rng.standard_normal((n, int(gauss.sum())))
        t["mu"] = 0.5 * rng.standard_normal((n, L))
        t["x"] = np.ones((n, 1, N))
        t["w"] = np.zeros((n, L))
        t["v"] = np.zeros((n, L))
    ref.preprocess.fill_params(params)
    ref.preprocess.fill_trials(trials)
    ref.gp.make_cholesky(trials, params, config)
    ref.core.update_w(trials, params, config)
    ref.core.update_v(trials, params, config)
    if window:
        config["window"] = window
        np.random.seed(seed)
        segs = ref.util.cut_trials(trials, params, config)
        segs = [copy.deepcopy(s) for s in segs]     # de-alias (hard part 3)
        ref.gp.make_cholesky(segs, params, config)
        ref.preprocess.fill_trials(segs)
        return segs, params, config
    return trials, params, config


def pack_state(prefix, trials, params):
    d = {}
    d[prefix + "mu"] = np.stack([t["mu"] for t in trials])
    d[prefix + "v"] = np.stack([t["v"] for t in trials])
    d[prefix + "w"] = np.stack([t["w"] for t in trials])
    d[prefix + "dmu"] = np.stack([t["dmu"] for t in trials])
    for k in ("a", "b", "noise", "omega", "sigma", "da", "db"):
        d[prefix + k] = np.array(params[k])
    return d


def golden_ichol(ref):
    out = {}
    for n in (50, 200, 500, 1000, 2000):
        for om in (5e-4, 5e-3, 5e-2):
            G = ref.math.ichol_gauss(n, om, 50)
            G2, piv = orc.ichol_gauss(n, om, 50, return_pivots=True)
            assert np.array_equal(G, G2), "oracle ichol differs bitwise from reference at n=%d omega=%g" % (n, om)
            key = "n%d_w%g" % (n, om)
            out[key + "_piv"] = piv
            out[key + "_ncol"] = np.array(int((np.abs(G).sum(axis=0) > 0).sum()))
            if n <= 200:
                out[key + "_G"] = G
            else:
                out[key + "_Grows"] = G[::25]           # every 25th row
                out[key + "_colsum"] = G.sum(axis=0)
                out[key + "_diagK"] = np.sum(G * G, axis=1)
    # the reference's own known-answer test (tests/test_math.py:7-14) at a smaller n: full-rank factor reproduces K
    G = ref.math.ichol_gauss(60, 1.0, 60)
    out["fullrank_n60_G"] = G
    np.savez_compressed(os.path.join(OUT, "ichol.npz"), **out)
    print("ichol.npz", len(out))


def golden_estep(ref):
    out = {}
    cases = {
        "poisson": dict(seed=1, n_trials=3, T=100, N=12, L=2, lik=None),
        "mixed": dict(seed=2, n_trials=2, T=100, N=9, L=3, lik=["poisson"] * 5 + ["gaussian"] * 4),
    }
    for name, kw in cases.items():
        for niter in (1, 25):
            segs, params, config = build_problem(ref, window=50, Eniter=niter, **kw)
            p = "%s_it%d_" % (name, niter)
            out[p + "y"] = np.stack([s["y"] for s in segs])
            out[p + "poisson"] = np.asarray(params["likelihood"]) == "poisson"
            out[p + "G"] = params["cholesky"][50]
            out.update(pack_state(p + "in_", segs, params))
            ref.core.estep(segs, params, config)
            out.update(pack_state(p + "out_", segs, params))
    # MAP variant (method != "VB": v is left untouched, core.py:105)
    segs, params, config = build_problem(ref, window=50, Eniter=3, method="MAP", **cases["poisson"])
    p = "map_it3_"
    out[p + "y"] = np.stack([s["y"] for s in segs])
    out[p + "poisson"] = np.asarray(params["likelihood"]) == "poisson"
    out[p + "G"] = params["cholesky"][50]
    out.update(pack_state(p + "in_", segs, params))
    ref.core.estep(segs, params, config)
    out.update(pack_state(p + "out_", segs, params))
    np.savez_compressed(os.path.join(OUT, "estep.npz"), **out)
    print("estep.npz", len(out))


def golden_mstep(ref):
    out = {}
    cases = {
        "poisson": dict(seed=3, n_trials=3, T=100, N=12, L=2, lik=None),
        "mixed": dict(seed=4, n_trials=2, T=100, N=9, L=3, lik=["poisson"] * 5 + ["gaussian"] * 4),
    }
    for name, kw in cases.items():
        for niter in (1, 25):
            segs, params, config = build_problem(ref, window=50, Eniter=2, Mniter=niter, **kw)
            ref.core.estep(segs, params, config)    # realistic v, w
            p = "%s_it%d_" % (name, niter)
            out[p + "y"] = np.stack([s["y"] for s in segs])
            out[p + "poisson"] = np.asarray(params["likelihood"]) == "poisson"
            out.update(pack_state(p + "in_", segs, params))
            ref.core.mstep(segs, params, config)
            out.update(pack_state(p + "out_", segs, params))
    np.savez_compressed(os.path.join(OUT, "mstep.npz"), **out)
    print("mstep.npz", len(out))


def golden_hstep(ref):
    out = {}
    segs, params, config = build_problem(ref, seed=5, n_trials=4, T=100, N=12, L=2, window=50, Eniter=5)
    ref.core.estep(segs, params, config)
    mu = np.stack([s["mu"] for s in segs])
    w = np.stack([s["w"] for s in segs])
    out["mu"], out["w"] = mu, w
    t = np.arange(50) * 1.0
    omegas = np.array([1e-3, 7e-3, 4e-2])
    out["omegas"] = omegas
    mask = np.array([0, 1, 0])
    for l in range(2):
        vals, grads = [], []
        for om in omegas:
            hyper = np.array([1.0, om, 1e-4])
            S = ref.gp.construct_posterior_cov(t, w[:, :, l].T, hyper)
            ll, dll = ref.gp.elbo(hyper, mask, t, mu[:, :, l].T, S)
            vals.append(ll)
            grads.append(dll)
        out["ll_l%d" % l] = np.array(vals)
        out["dll_l%d" % l] = np.array(grads)
        initial = (1.0, params["omega"][l], 1e-4)
        bounds = ((1e-3, 1), config["omega_bound"], (1e-4 / 2, 1e-4 * 2))
        opt, fun = ref.gp.optimze1d(t, mu[:, :, l].T, w[:, :, l].T, initial, bounds, mask=mask)
        out["opt_l%d" % l] = np.asarray(opt)
        out["fun_l%d" % l] = np.asarray(fun)
    out["omega0"] = np.array(params["omega"])
    # whole hstep (gp.optimize) on the same segments
    ref.core.hstep(segs, params, config)
    out["omega_after"] = np.array(params["omega"])
    out["sigma_after"] = np.array(params["sigma"])
    out["G_after"] = params["cholesky"][50]
    np.savez_compressed(os.path.join(OUT, "hstep.npz"), **out)
    print("hstep.npz", len(out))


def golden_update_wv(ref):
    out = {}
    trials, params, config = build_problem(ref, seed=6, n_trials=2, T=1000, N=12, L=2)
    out["y"] = np.stack([t["y"] for t in trials]).astype(np.uint8)
    assert np.array_equal(out["y"], np.stack([t["y"] for t in trials]))
    # build_problem already ran update_w / update_v once from v = 0: run a second round so v enters w
    for k in ("mu",):
        out["in_" + k] = np.stack([t[k] for t in trials])
    out["in_v"] = np.stack([t["v"] for t in trials])
    for k in ("a", "b", "noise", "omega", "sigma"):
        out[k] = np.array(params[k])
    ref.core.update_w(trials, params, config)
    ref.core.update_v(trials, params, config)
    out["out_w"] = np.stack([t["w"] for t in trials])
    out["out_v"] = np.stack([t["v"] for t in trials])
    # a short full-trial E-step (the final ``infer`` regime: T x 50 factor)
    config["Eniter"] = 3
    ref.core.estep(trials, params, config)
    out["infer_mu"] = np.stack([t["mu"] for t in trials])
    out["infer_v"] = np.stack([t["v"] for t in trials])
    out["infer_w"] = np.stack([t["w"] for t in trials])
    out["infer_dmu"] = np.stack([t["dmu"] for t in trials])
    np.savez_compressed(os.path.join(OUT, "update_wv.npz"), **out)
    print("update_wv.npz", len(out))


def golden_vem(ref):
    out = {}
    segs, params, config = build_problem(ref, seed=7, n_trials=4, T=100, N=10, L=2, window=50,
                                         max_iter=3, min_iter=3)
    out["y"] = np.stack([s["y"] for s in segs])
    out.update(pack_state("in_", segs, params))
    ref.core.vem(segs, params, config)
    out.update(pack_state("out_", segs, params))
    out["n_it"] = np.array(config["runtime"]["it"])
    np.savez_compressed(os.path.join(OUT, "vem.npz"), **out)
    print("vem.npz", len(out))


def golden_fit(ref):
    out = {}
    trials = make_trials(10, 200, 30, 3, seed=0)
    out["y"] = np.stack([t["y"] for t in trials]).astype(np.uint8)
    np.random.seed(0)
    traj = {"omega": [], "sigma": []}

    def record(trials_, params_, config_):      # vem calls this after every H-step (vlgp/core.py:339-343)
        traj["omega"].append(np.array(params_["omega"], dtype=float))
        traj["sigma"].append(np.array(params_["sigma"], dtype=float))

    res = ref.fit(trials, 3, max_iter=3, min_iter=3, callbacks=[record])
    out["mu"] = np.stack([t["mu"] for t in res["trials"]])
    out["v"] = np.stack([t["v"] for t in res["trials"]])
    out["w"] = np.stack([t["w"] for t in res["trials"]])
    for k in ("a", "b", "noise", "omega", "sigma"):
        out[k] = np.array(res["params"][k])
    for k in ("a", "b", "omega"):
        out["initial_" + k] = np.array(res["params"]["initial"][k])
    # omega / sigma as the reference's H-step left them after each EM iteration: lets a test run the default fit with
    # this trajectory injected, so that everything around the optimiser is compared without the pivot-tie sensitivity
    out["omega_traj"] = np.stack(traj["omega"])
    out["sigma_traj"] = np.stack(traj["sigma"])
    np.savez_compressed(os.path.join(OUT, "fit_tutorial.npz"), **out)
    print("fit_tutorial.npz", len(out))


def golden_fit_wide_window(ref):
    """fit() with window=100 (the H-step's wide path, csrc/hstep_wide.cu): outputs and the omega of every iteration."""
    out = {}
    trials = make_trials(4, 200, 12, 2, seed=5)
    out["y"] = np.stack([t["y"] for t in trials]).astype(np.uint8)
    np.random.seed(0)
    traj = {"omega": [], "sigma": []}

    def record(trials_, params_, config_):
        traj["omega"].append(np.array(params_["omega"], dtype=float))
        traj["sigma"].append(np.array(params_["sigma"], dtype=float))

    res = ref.fit(trials, 2, window=100, max_iter=3, min_iter=3, callbacks=[record])
    for k in ("mu", "v", "w"):
        out[k] = np.stack([t[k] for t in res["trials"]])
    for k in ("a", "b", "noise", "omega", "sigma"):
        out[k] = np.array(res["params"][k])
    out["omega_traj"] = np.stack(traj["omega"])
    out["sigma_traj"] = np.stack(traj["sigma"])
    np.savez_compressed(os.path.join(OUT, "fit_wide_window.npz"), **out)
    print("fit_wide_window.npz", len(out))


def golden_fit_fixed_omega(ref):
    """fit() with Hstep=False: omega stays at its initial value, so the prior factors (and their pivot sets) are the
    same on both sides and the whole chain initialise -> update_w/v -> cut -> vem -> infer can be compared tightly."""
    out = {}
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = ref.fit(trials, 3, max_iter=3, min_iter=3, Hstep=False)
    out["mu"] = np.stack([t["mu"] for t in res["trials"]])
    out["v"] = np.stack([t["v"] for t in res["trials"]])
    out["w"] = np.stack([t["w"] for t in res["trials"]])
    for k in ("a", "b", "noise", "omega", "sigma"):
        out[k] = np.array(res["params"][k])
    np.savez_compressed(os.path.join(OUT, "fit_fixed_omega.npz"), **out)
    print("fit_fixed_omega.npz", len(out))


def golden_api_extras(ref):
    """After fit(Hstep=False) on the tutorial-shaped problem: sample_posterior of the first trial (vlgp/api.py:142-168)
    and transform of three NEW trials with the fitted model (vlgp/api.py:171-184)."""
    out = {}
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = ref.fit(trials, 3, max_iter=3, min_iter=3, Hstep=False)
    params, config = res["params"], res["config"]
    np.random.seed(5)
    out["samples"] = ref.sample_posterior(res["trials"][0], params, 4)
    out["trial0_mu"] = res["trials"][0]["mu"].copy()
    out["trial0_w"] = res["trials"][0]["w"].copy()
    for k in ("omega", "sigma"):
        out[k] = np.array(params[k])
    new = make_trials(3, 200, 30, 3, seed=77)
    out["new_y"] = np.stack([t["y"] for t in new]).astype(np.uint8)
    ref.transform(new, params, config)
    for k in ("mu", "v", "w"):
        out["new_" + k] = np.stack([t[k] for t in new])
    np.savez_compressed(os.path.join(OUT, "api_extras.npz"), **out)
    print("api_extras.npz", len(out))


def fit_option_cases():
    """kwargs of fit() beyond the defaults (vlgp/preprocess.py:59-74,85-106); shared with tests/test_host_orchestration.py."""
    rng = np.random.default_rng(5)
    N, L = 12, 2
    return N, L, {
        "mixed_lik": dict(max_iter=2, min_iter=2, lik=["poisson"] * 8 + ["gaussian"] * 4),
        "user_a_b": dict(max_iter=2, min_iter=2, a=0.4 * rng.standard_normal((L, N)), b=np.full((1, N), -2.0)),
        "user_omega_sigma": dict(max_iter=2, min_iter=2, omega=np.array([0.01, 0.02]), sigma=np.array([0.8, 1.0])),
        "latent_both_map": dict(max_iter=2, min_iter=2, constrain_latent="both", method="MAP"),
        "loading_svd": dict(max_iter=2, min_iter=2, constrain_loading="svd"),
        "tol_early_stop": dict(max_iter=6, min_iter=1, tol=1e-1),
        "window25": dict(max_iter=2, min_iter=2, window=25),
    }


def fit_option_trials(N, L, kw):
    trials = make_trials(4, 100, N, L, seed=9)
    if "lik" in kw:
        rng = np.random.default_rng(1)
        for t in trials:
            t["y"][:, 8:] = t["y"][:, 8:] + 0.3 * rng.standard_normal((100, 4))
    return trials


def golden_fit_options(ref):
    """Whole fit() runs of the reference under non-default keyword arguments (small problem, T % window == 0)."""
    out = {}
    N, L, cases = fit_option_cases()
    for name, kw in cases.items():
        trials = fit_option_trials(N, L, kw)
        np.random.seed(0)
        res = ref.fit(trials, L, **copy.deepcopy(kw))
        p = name + "/"
        for k in ("mu", "v", "w"):
            out[p + k] = np.stack([t[k] for t in res["trials"]])
        for k in ("a", "b", "noise", "omega", "sigma"):
            out[p + k] = np.array(res["params"][k])
        out[p + "n_it"] = np.array(res["config"]["runtime"]["it"])
    np.savez_compressed(os.path.join(OUT, "fit_options.npz"), **out)
    print("fit_options.npz", len(out))


def fit_overlap_trials():
    """Unequal trial lengths that are not multiples of the window: overlapping, ALIASED segments (vlgp/util.py:482-498)."""
    out = []
    for i, T in enumerate((130, 175, 100, 262)):
        out += make_trials(1, T, 14, 2, seed=40 + i)
    return out


FIT_OVERLAP_CASES = {
    "default": dict(max_iter=3, min_iter=3),
    "latent_both_no_hstep": dict(max_iter=3, min_iter=3, constrain_latent="both", Hstep=False),
    "row_norm_loading": dict(max_iter=2, min_iter=2, constrain_loading=2),
}


def golden_fit_overlap(ref):
    """Whole fit() of the reference on trials whose windows overlap, i.e. with its sequential, in-place updates of the
    bins two segments share."""
    out = {}
    for name, kw in FIT_OVERLAP_CASES.items():
        trials = fit_overlap_trials()
        np.random.seed(0)
        res = ref.fit(trials, 2, **copy.deepcopy(kw))
        p = name + "/"
        for k in ("mu", "v", "w"):
            out[p + k] = np.concatenate([t[k] for t in res["trials"]])
        for k in ("a", "b", "noise", "omega", "sigma"):
            out[p + k] = np.array(res["params"][k])
    np.savez_compressed(os.path.join(OUT, "fit_overlap.npz"), **out)
    print("fit_overlap.npz", len(out))


VEM_OPTION_CASES = {
    # name: (likelihood list or None, config overrides) -- the option branches of vem that the default fit never takes
    "latent_both": (None, dict(constrain_latent="both")),
    "loading_svd": (None, dict(constrain_loading="svd")),
    "loading_row2_latent_location": (None, dict(constrain_loading=2, constrain_latent="location")),
    "latent_scale_no_loading": (None, dict(constrain_loading="none", constrain_latent="scale")),
    "gradient_step": (None, dict(use_hessian=False, learning_rate=1e-4)),
    "map_no_hstep": (None, dict(method="MAP", Hstep=False)),
    "mixed_lik": (["poisson"] * 6 + ["gaussian"] * 4, dict()),
    "short_steps_tight_bounds": (None, dict(Eniter=3, Mniter=2, dmu_bound=0.05, da_bound=0.01, db_bound=0.02)),
}


def golden_vem_options(ref):
    """Two vem iterations of the reference under the option branches listed above (vlgp/core.py:191-235,366-416)."""
    out = {}
    for name, (lik, kw) in VEM_OPTION_CASES.items():
        segs, params, config = build_problem(ref, seed=11, n_trials=4, T=100, N=10, L=2, lik=lik, window=50,
                                             max_iter=2, min_iter=2, **kw)
        p = name + "/"
        out[p + "y"] = np.stack([s["y"] for s in segs])
        out[p + "poisson"] = params["likelihood"] == "poisson"
        out.update(pack_state(p + "in_", segs, params))
        ref.core.vem(segs, params, config)
        out.update(pack_state(p + "out_", segs, params))
        out[p + "n_it"] = np.array(config["runtime"]["it"])
    np.savez_compressed(os.path.join(OUT, "vem_options.npz"), **out)
    print("vem_options.npz", len(out))


VEM_REGRESSOR_CASES = {
    # name: (likelihoods or None, number of regressors xdim = max(history, 1), config keywords)
    "history2_poisson": (None, 3, dict(Hstep=False)),
    "history1_mixed_hstep": (["poisson"] * 6 + ["gaussian"] * 4, 2, dict()),
    "scaled_bias_only": (None, 1, dict(Hstep=False, use_hessian=False, learning_rate=1e-4)),
}


def regressor_design(y, xdim, scale_bias=1.0):
    """x (bins, xdim, neurons): the bias column (times scale_bias) and xdim - 1 spike-history regressors y[t - k]."""
    T, N = y.shape
    x = np.zeros((T, xdim, N))
    x[:, 0, :] = scale_bias
    for k in range(1, xdim):
        x[k:, k, :] = y[:-k]
    return x


def golden_vem_regressors(ref):
    """Two vem iterations of the reference with regressors other than the all-ones bias column (vlgp/core.py:66,
    205-220,229-235): spike-history designs with xdim = 3 (Poisson) and xdim = 2 (mixed likelihoods, H-step on), and a
    single regressor that is not all ones.  The reference takes b of shape (xdim, N) from the caller in that case."""
    out = {}
    for name, (lik, xdim, kw) in VEM_REGRESSOR_CASES.items():
        segs, params, config = build_problem(ref, seed=23, n_trials=4, T=100, N=10, L=2, lik=lik, window=50,
                                             max_iter=2, min_iter=2, **kw)
        rng = np.random.default_rng(5)
        b = np.zeros((xdim, 10))
        b[0] = params["b"][0] / (0.5 if name == "scaled_bias_only" else 1.0)
        b[1:] = 0.05 * rng.standard_normal((xdim - 1, 10))
        params["b"], params["db"], params["xdim"] = b, np.zeros_like(b), xdim
        for sg in segs:
            sg["x"] = regressor_design(sg["y"], xdim, 0.5 if name == "scaled_bias_only" else 1.0)
        ref.core.update_w(segs, params, config)
        ref.core.update_v(segs, params, config)
        p = name + "/"
        out[p + "y"] = np.stack([s["y"] for s in segs])
        out[p + "poisson"] = params["likelihood"] == "poisson"
        out.update(pack_state(p + "in_", segs, params))
        ref.core.vem(segs, params, config)
        out.update(pack_state(p + "out_", segs, params))
        out[p + "n_it"] = np.array(config["runtime"]["it"])
    np.savez_compressed(os.path.join(OUT, "vem_regressors.npz"), **out)
    print("vem_regressors.npz", len(out))


def golden_gpfa(ref):
    """vlgp/gpfa.py: em() on stacked segments -- identity noise (what prepare() passes) and a non-uniform initial R (which
    exposes the reference's time-major bigR against its neuron-major bigC) -- and sekernel."""
    import vlgp.gpfa as rgpfa
    from vlgp.gp import sekernel

    rng = np.random.default_rng(21)
    out = {}
    for case, (m, n, ydim, zdim, iters, uniform) in {"eye": (12, 20, 8, 2, 3, True), "noise": (9, 25, 7, 3, 4, False),
                                                     "one": (6, 50, 12, 2, 1, True)}.items():
        y = rng.poisson(1.0, (m, n, ydim)).astype(float)
        C = rng.standard_normal((zdim, ydim))
        d = y.mean(axis=(0, 1))[None, :]
        R = np.eye(ydim) if uniform else np.diag(0.5 + rng.random(ydim))
        K = sekernel(np.arange(n) * 1.0, 1.0, 3.0)
        z, C2, d2, R2 = rgpfa.em(y.copy(), C.copy(), d.copy(), R.copy(), K, iters)
        pre = case + "_"
        out.update({pre + "y": y, pre + "C": C, pre + "d": d, pre + "R": R, pre + "K": K, pre + "iters": iters,
                    pre + "out_z": z, pre + "out_C": C2, pre + "out_d": d2, pre + "out_R": R2})
    np.savez_compressed(os.path.join(OUT, "gpfa.npz"), **out)
    print("gpfa.npz", len(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    import vlgp.preprocess, vlgp.core, vlgp.gp, vlgp.math, vlgp.util  # noqa: F401,E401
    only = sys.argv[1:]
    for fn in (golden_ichol, golden_estep, golden_mstep, golden_hstep, golden_update_wv, golden_vem, golden_fit,
               golden_fit_fixed_omega, golden_fit_wide_window, golden_vem_options, golden_api_extras,
               golden_fit_options, golden_fit_overlap, golden_vem_regressors, golden_gpfa):
        if not only or fn.__name__.replace("golden_", "") in only:
            fn(ref)


if __name__ == "__main__":
    main()
