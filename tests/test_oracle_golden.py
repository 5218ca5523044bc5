"""Pin the NumPy oracle (oracle/vlgp_oracle.py) to the golden vectors made from the unmodified reference
(oracle/make_golden.py).  CPU only.  Tolerance: 1e-11 relative -- the oracle restates the same formulas with the
same LAPACK drivers, so differences are a few ulps amplified by at most 25 Newton iterations."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import vlgp_oracle as orc

TOL = 1e-11


def _segs(g, p, names=("mu", "v", "w", "dmu")):
    n = g[p + "in_mu"].shape[0]
    N = g[p + "y"].shape[2]
    return [dict(y=g[p + "y"][i], x=np.ones((g[p + "y"].shape[1], 1, N)),
                 **{k: g[p + "in_" + k][i].copy() for k in names}) for i in range(n)]


def _params(g, p, lik_key=None):
    poisson = g[(lik_key or p) + "poisson"]
    a = g[p + "in_a"].copy()
    return dict(a=a, b=g[p + "in_b"].copy(), noise=g[p + "in_noise"].copy(), omega=g[p + "in_omega"].copy(),
                sigma=g[p + "in_sigma"].copy(), da=np.zeros_like(a), db=np.zeros_like(g[p + "in_b"]),
                likelihood=np.where(poisson, "poisson", "gaussian"), zdim=a.shape[0], ydim=a.shape[1], xdim=1,
                rank=50, gp_noise=1e-4, dt=1)


def test_ichol_matches_reference_bitwise():
    g = load_golden("ichol")
    for n in (50, 200, 500, 1000, 2000):
        for om in (5e-4, 5e-3, 5e-2):
            key = "n%d_w%g" % (n, om)
            G, piv = orc.ichol_gauss(n, om, 50, return_pivots=True)
            assert np.array_equal(piv, g[key + "_piv"])
            assert int((np.abs(G).sum(axis=0) > 0).sum()) == int(g[key + "_ncol"])
            if n <= 200:
                assert np.array_equal(G, g[key + "_G"])
            else:
                assert np.array_equal(G[::25], g[key + "_Grows"])
                assert relerr(G.sum(axis=0), g[key + "_colsum"]) < 1e-13


def test_ichol_known_answer_full_rank():
    """The reference's own known-answer test (tests/test_math.py:7-14): a full-rank factor reproduces K."""
    n, omega = 60, 1.0
    K = np.exp(-omega * (np.arange(n)[:, None] - np.arange(n)[None, :]) ** 2)
    G = orc.ichol_gauss(n, omega, n)
    assert np.allclose(K, G @ G.T)
    assert np.array_equal(G, load_golden("ichol")["fullrank_n60_G"])
    n = 500
    K = np.exp(-omega * (np.arange(n)[:, None] - np.arange(n)[None, :]) ** 2)
    G = orc.ichol_gauss(n, omega, n)
    assert np.allclose(K, G @ G.T)


@pytest.mark.parametrize("case", ["poisson_it1", "poisson_it25", "mixed_it1", "mixed_it25", "map_it3"])
def test_estep(case):
    g = load_golden("estep")
    p = case + "_"
    segs = _segs(g, p)
    params = _params(g, p)
    params["cholesky"] = {50: g[p + "G"]}
    cfg = orc.default_config(Eniter=int(case.split("it")[1]), method="MAP" if case.startswith("map") else "VB")
    orc.estep(segs, params, cfg)
    for k in ("mu", "v", "w", "dmu"):
        got = np.stack([s[k] for s in segs])
        # dmu shrinks to ~1e-10 as the Newton iterations converge: measure it on the scale of mu
        scale = g[p + "out_mu"] if k == "dmu" else g[p + "out_" + k]
        assert np.max(np.abs(got - g[p + "out_" + k])) <= TOL * np.max(np.abs(scale)), k


@pytest.mark.parametrize("case", ["poisson_it1", "poisson_it25", "mixed_it1", "mixed_it25"])
def test_mstep(case):
    g = load_golden("mstep")
    p = case + "_"
    segs = _segs(g, p)
    params = _params(g, p)
    cfg = orc.default_config(Mniter=int(case.split("it")[1]))
    orc.mstep(segs, params, cfg)
    for k in ("a", "b", "noise", "da", "db"):
        assert relerr(params[k], g[p + "out_" + k]) < TOL, k


def test_hstep_objective_and_optimum():
    g = load_golden("hstep")
    t = np.arange(50) * 1.0
    for l in range(2):
        mu, w = g["mu"][:, :, l].T, g["w"][:, :, l].T
        for i, om in enumerate(g["omegas"]):
            hyper = np.array([1.0, om, 1e-4])
            S = orc.posterior_cov(t, w, hyper)
            ll, dll = orc.elbo(hyper, t, mu, S)
            assert abs(ll - g["ll_l%d" % l][i]) < 1e-10 * abs(g["ll_l%d" % l][i])
            assert abs(dll - g["dll_l%d" % l][i][1]) < 1e-8 * max(abs(g["dll_l%d" % l][i][1]), 1.0)
            assert g["dll_l%d" % l][i][0] == 0 and g["dll_l%d" % l][i][2] == 0


def test_hstep_whole():
    g = load_golden("hstep")
    n = g["mu"].shape[0]
    segs = [dict(y=np.zeros((50, 1)), mu=g["mu"][i], w=g["w"][i]) for i in range(n)]
    params = dict(omega=g["omega0"].copy(), sigma=np.ones(2), gp_noise=1e-4, dt=1, zdim=2, rank=50)
    cfg = orc.default_config()
    orc.hstep(segs, params, cfg)
    assert relerr(params["omega"], g["omega_after"]) < 1e-6
    assert relerr(params["sigma"], g["sigma_after"]) < 1e-12
    assert relerr(params["cholesky"][50] @ params["cholesky"][50].transpose(0, 2, 1),
                  g["G_after"] @ g["G_after"].transpose(0, 2, 1)) < 1e-5


def test_update_w_v_and_long_trial_estep():
    g = load_golden("update_wv")
    y = g["y"].astype(float)
    N = y.shape[2]
    trials = [dict(y=y[i], x=np.ones((y.shape[1], 1, N)), mu=g["in_mu"][i].copy(), v=g["in_v"][i].copy())
              for i in range(y.shape[0])]
    L = g["a"].shape[0]
    params = dict(a=g["a"], b=g["b"], noise=g["noise"], omega=g["omega"], sigma=g["sigma"], zdim=L, rank=50,
                  likelihood=np.array(["poisson"] * N))
    params["cholesky"] = orc.make_cholesky([y.shape[1]], params["omega"], params["sigma"], 50)
    cfg = orc.default_config(Eniter=3)
    orc.update_w(trials, params, cfg)
    orc.update_v(trials, params, cfg)
    assert relerr(np.stack([t["w"] for t in trials]), g["out_w"]) < TOL
    assert relerr(np.stack([t["v"] for t in trials]), g["out_v"]) < TOL
    for t in trials:
        t["dmu"] = np.zeros_like(t["mu"])
    orc.estep(trials, params, cfg)
    for k in ("mu", "v", "w", "dmu"):
        scale = g["infer_mu"] if k == "dmu" else g["infer_" + k]
        assert np.max(np.abs(np.stack([t[k] for t in trials]) - g["infer_" + k])) < TOL * np.max(np.abs(scale)), k


def test_vem_three_iterations():
    g = load_golden("vem")
    segs = _segs(g, "")
    N = g["y"].shape[2]
    params = _params({**g, "poisson": np.ones(N, bool)}, "")
    params["cholesky"] = orc.make_cholesky([50], params["omega"], params["sigma"], 50)
    cfg = orc.default_config(max_iter=3, min_iter=3)
    orc.vem(segs, params, cfg)
    assert cfg["runtime"]["it"] == int(g["n_it"])
    assert relerr(params["omega"], g["out_omega"]) < 1e-6
    for k in ("a", "b"):
        assert relerr(params[k], g["out_" + k]) < 1e-7, k
    assert relerr(np.stack([s["mu"] for s in segs]), g["out_mu"]) < 1e-6


def test_cut_trial_starts_matches_reference_rule():
    np.random.seed(3)
    s = orc.cut_trial_starts(230, 50)
    assert len(s) == 5 and s[0] == 0 and s[-1] == 180 and np.all(np.diff(s) <= 50)
    np.random.seed(3)
    s2 = orc.cut_trial_starts(200, 50)
    assert np.array_equal(s2, [0, 50, 100, 150])


def test_fit_fixed_omega_pipeline():
    """The whole fit() pipeline restated with the oracle (host set-up from vlgp_b200.preprocess/util, which mirror the
    reference's RNG use) against the reference's fit(Hstep=False) golden: pins initialise + cut_trials + vem + infer."""
    from vlgp_b200 import preprocess
    from vlgp_b200.synth import make_trials
    from vlgp_b200.util import cut_trials

    g = load_golden("fit_fixed_omega")
    trials = make_trials(10, 200, 30, 3, seed=0)
    config = preprocess.get_config(max_iter=3, min_iter=3, Hstep=False)
    params = preprocess.get_params(trials, 3, omega_bound=config["omega_bound"])
    np.random.seed(0)
    preprocess.initialize(trials, params, config)
    preprocess.fill_params(params)
    preprocess.fill_trials(trials)
    params["cholesky"] = orc.make_cholesky([200], params["omega"], params["sigma"], 50)
    orc.update_w(trials, params)
    orc.update_v(trials, params, config)
    segs = list(cut_trials(trials, params, config))
    preprocess.fill_trials(segs)
    params["cholesky"] = orc.make_cholesky([50], params["omega"], params["sigma"], 50)
    orc.vem(segs, params, config)
    params["cholesky"] = orc.make_cholesky([200], params["omega"], params["sigma"], 50)
    orc.update_w(trials, params)
    orc.update_v(trials, params, config)
    orc.estep(trials, params, config, n_iter=config["max_iter"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in trials]), g[k]) < 1e-9, k
    for k in ("a", "b", "noise"):
        assert relerr(params[k], g[k]) < 1e-9, k


VEM_OPTION_CASES = {
    "latent_both": dict(constrain_latent="both"),
    "loading_svd": dict(constrain_loading="svd"),
    "loading_row2_latent_location": dict(constrain_loading=2, constrain_latent="location"),
    "latent_scale_no_loading": dict(constrain_loading="none", constrain_latent="scale"),
    "gradient_step": dict(use_hessian=False, learning_rate=1e-4),
    "map_no_hstep": dict(method="MAP", Hstep=False),
    "mixed_lik": dict(),
    "short_steps_tight_bounds": dict(Eniter=3, Mniter=2, dmu_bound=0.05, da_bound=0.01, db_bound=0.02),
}


@pytest.mark.parametrize("case", sorted(VEM_OPTION_CASES))
def test_vem_option_branches(case):
    """Two vem iterations of the reference under the options the default fit never takes: latent / loading
    constraints (vlgp/core.py:366-416), gradient-step M-step (:194-198), MAP (no variance update), Gaussian channels,
    active clipping bounds.  Same configuration table as oracle/make_golden.py::VEM_OPTION_CASES."""
    g = load_golden("vem_options")
    p = case + "/"
    segs = _segs(g, p)
    params = _params(g, p)
    params["cholesky"] = orc.make_cholesky([50], params["omega"], params["sigma"], 50)
    cfg = orc.default_config(max_iter=2, min_iter=2, **VEM_OPTION_CASES[case])
    orc.vem(segs, params, cfg)
    assert cfg["runtime"]["it"] == int(g[p + "n_it"])
    hstep = cfg["Hstep"]
    assert relerr(params["omega"], g[p + "out_omega"]) < (1e-6 if hstep else 1e-15)
    tol = 1e-6 if hstep else 1e-9        # with the H-step the new prior factor depends on omega's last digits
    for k in ("a", "b", "noise"):
        assert relerr(params[k], g[p + "out_" + k]) < tol, k
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([s[k] for s in segs]), g[p + "out_" + k]) < tol, k


VEM_REGRESSOR_CASES = {
    "history2_poisson": (3, 0.0, dict(Hstep=False)),
    "history1_mixed_hstep": (2, 0.0, dict()),
    "scaled_bias_only": (1, 0.5, dict(Hstep=False, use_hessian=False, learning_rate=1e-4)),
}


def regressor_design(y, xdim, scale_bias):
    """Same design as oracle/make_golden.py::regressor_design: bias column and xdim - 1 spike-history regressors."""
    T, N = y.shape
    x = np.zeros((T, xdim, N))
    x[:, 0, :] = scale_bias or 1.0
    for k in range(1, xdim):
        x[k:, k, :] = y[:-k]
    return x


@pytest.mark.parametrize("case", sorted(VEM_REGRESSOR_CASES))
def test_vem_general_regressors(case):
    """Two vem iterations of the reference with regressors other than the all-ones bias column (vlgp/core.py:66,
    205-220,229-235): xdim = 3 spike-history design, xdim = 2 with Gaussian channels and the H-step, one scaled column."""
    g = load_golden("vem_regressors")
    p = case + "/"
    xdim, scale, kw = VEM_REGRESSOR_CASES[case]
    segs = _segs(g, p)
    for sg in segs:
        sg["x"] = regressor_design(sg["y"], xdim, scale)
    params = _params(g, p)
    params["xdim"] = xdim
    params["cholesky"] = orc.make_cholesky([50], params["omega"], params["sigma"], 50)
    cfg = orc.default_config(max_iter=2, min_iter=2, **kw)
    orc.vem(segs, params, cfg)
    assert params["b"].shape == (xdim, 10)
    tol = 1e-6 if cfg["Hstep"] else 1e-9
    for k in ("a", "b", "noise"):
        assert relerr(params[k], g[p + "out_" + k]) < tol, k
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([s[k] for s in segs]), g[p + "out_" + k]) < tol, k


def test_posterior_cov_low_rank_identity_matches_the_reference_expression():
    """The device computes the sample_posterior covariance inv(inv(K + reg I) + W) (vlgp/api.py:160-166) through
    reg E^-1 + F Q^-1 F' (csrc/postcov.cu).  Here the identity itself is pinned, in NumPy, against the reference's
    expression and against the covariance behind the reference's golden samples (tests/golden/api_extras.npz): equal to
    the reference's own asymmetry (1e-9); reg = 0 reproduces util.posterior_cov (vlgp/util.py:541-547)."""
    g = load_golden("api_extras")
    G = orc.make_cholesky([200], g["omega"], g["sigma"], 50)[200]
    w = g["trial0_w"]
    for k in range(3):
        K = G[k] @ G[k].T
        for reg in (1e-6, 1e-3):
            ref = np.linalg.inv(np.linalg.inv(K + reg * np.eye(200)) + np.diag(w[:, k]))
            E = 1.0 + reg * w[:, k]
            F = G[k] / E[:, None]
            Q = np.eye(50) + G[k].T @ ((w[:, k] / E)[:, None] * G[k])
            cov = np.diag(reg / E) + F @ np.linalg.solve(Q, F.T)
            assert relerr(cov, ref) < 5e-9
        woodbury = K - K @ np.linalg.solve(np.diag(1.0 / w[:, k]) + K, K)
        Q0 = np.eye(50) + G[k].T @ (w[:, k][:, None] * G[k])
        assert relerr(G[k] @ np.linalg.solve(Q0, G[k].T), woodbury) < 1e-9


def test_transform_new_trials_pipeline():
    """transform() on trials the model has not seen (vlgp/api.py:171-184): initialise through the fitted
    FactorAnalysis map, then max_iter E-step iterations on the uncut trials -- restated with the host set-up of
    vlgp_b200.preprocess and the oracle's E-step, against the reference's output."""
    from vlgp_b200 import preprocess
    from vlgp_b200.synth import make_trials
    from vlgp_b200.util import cut_trials

    g = load_golden("api_extras")
    trials = make_trials(10, 200, 30, 3, seed=0)
    config = preprocess.get_config(max_iter=3, min_iter=3, Hstep=False)
    params = preprocess.get_params(trials, 3, omega_bound=config["omega_bound"])
    np.random.seed(0)
    preprocess.initialize(trials, params, config)
    preprocess.fill_params(params)
    preprocess.fill_trials(trials)
    params["cholesky"] = orc.make_cholesky([200], params["omega"], params["sigma"], 50)
    orc.update_w(trials, params)
    orc.update_v(trials, params, config)
    segs = list(cut_trials(trials, params, config))
    preprocess.fill_trials(segs)
    params["cholesky"] = orc.make_cholesky([50], params["omega"], params["sigma"], 50)
    orc.vem(segs, params, config)
    params["cholesky"] = orc.make_cholesky([200], params["omega"], params["sigma"], 50)
    # -- transform(new, params, config)
    new = make_trials(3, 200, 30, 3, seed=77)
    assert np.array_equal(np.stack([t["y"] for t in new]), g["new_y"])
    preprocess.initialize(new, params, config)
    preprocess.fill_trials(new)
    orc.estep(new, params, config, n_iter=config["max_iter"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in new]), g["new_" + k]) < 1e-9, k


@pytest.mark.parametrize("case", ["eye", "noise", "one"])
def test_gpfa_em(case):
    """GPFA branch (vlgp/gpfa.py:20-56): the oracle's Kronecker-free restatement against the reference's em() -- identity
    noise, a non-uniform initial R (the reference's bigR ordering), a single iteration."""
    g = load_golden("gpfa")
    p = case + "_"
    z, C, d, R = orc.gpfa_em(g[p + "y"], g[p + "C"], g[p + "d"], g[p + "R"], g[p + "K"], int(g[p + "iters"]))
    for got, key in ((z, "z"), (C, "C"), (d, "d"), (R, "R")):
        assert relerr(got, g[p + "out_" + key]) < 1e-11, key
    assert relerr(orc.sekernel(np.arange(g[p + "K"].shape[0]) * 1.0, 1.0, 3.0), g[p + "K"]) < 1e-15
