"""GPU parity tests: the CUDA path (through the ctypes C ABI) against the golden vectors of the unmodified reference
(tests/golden/*.npz, made by oracle/make_golden.py) and against the NumPy oracle on seeded inputs.

Tolerances (fp64 device arithmetic, stated per SURVEY.md section 8(c)):
  single step (E-step, M-step, update_w/v):  <= 1e-10 relative to the largest entry of the reference output
  H-step objective:                          ll <= 1e-9 relative, dll <= 1e-7 relative (K has condition number ~1e6)
  three vem iterations / full fit:           <= 1e-5 relative on posterior means (north-star tolerance)
"""
import copy

import numpy as np
import pytest

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-10


@pytest.fixture(scope="module")
def vl():
    import vlgp_b200
    from vlgp_b200 import core, gp, engine  # noqa: F401

    engine.get_engine()        # fail loudly here if the native library / GPU is missing
    return vlgp_b200


def _segs(g, p, names=("mu", "v", "w", "dmu")):
    n = g[p + "in_mu"].shape[0]
    N = g[p + "y"].shape[2]
    return [dict(y=g[p + "y"][i].astype(float), x=np.ones((g[p + "y"].shape[1], 1, N)),
                 **{k: g[p + "in_" + k][i].copy() for k in names}) for i in range(n)]


def _params(g, p, poisson=None):
    poisson = g[p + "poisson"] if poisson is None else poisson
    a = g[p + "in_a"].copy()
    return dict(a=a, b=g[p + "in_b"].copy(), noise=g[p + "in_noise"].copy(), omega=g[p + "in_omega"].copy(),
                sigma=g[p + "in_sigma"].copy(), da=np.zeros_like(a), db=np.zeros_like(g[p + "in_b"]),
                likelihood=np.where(poisson, "poisson", "gaussian"), zdim=a.shape[0], ydim=a.shape[1], xdim=1,
                rank=50, gp_noise=1e-4, dt=1)


def _cfg(**kw):
    from vlgp_b200.preprocess import get_config

    return get_config(**kw)


# ----------------------------------------------------------------------------------------------------------------------
def test_ichol_pivots_and_factor(vl):
    from vlgp_b200.gp import make_cholesky
    from oracle import vlgp_oracle as orc

    g = load_golden("ichol")
    omegas = np.array([5e-4, 5e-3, 5e-2])
    eng = __import__("vlgp_b200.engine", fromlist=["get_engine"]).get_engine()
    params = dict(ydim=3, zdim=3, xdim=1, rank=50, gp_noise=1e-4, dt=1, likelihood=np.array(["poisson"] * 3),
                  sigma=np.ones(3), omega=omegas)
    eng.ensure_model(params)
    eng.push_params(params, which=("sigma", "omega"))
    for n in (50, 200, 500, 1000, 2000):
        with eng.new_trials([n]) as ts:
            ts.make_cholesky()
            G, piv, ncol = ts.get_cholesky(n, with_pivots=True)
        for l, om in enumerate(omegas):
            key = "n%d_w%g" % (n, om)
            ref_piv = g[key + "_piv"]
            assert int(ncol[l]) == int(g[key + "_ncol"]) == len(ref_piv), key
            assert np.array_equal(piv[l, :len(ref_piv)], ref_piv), key
            assert np.all(piv[l, len(ref_piv):] == -1)
            Gref = orc.ichol_gauss(n, om, 50)
            assert relerr(G[l], Gref) < 1e-12, key
            if n <= 200:
                assert relerr(G[l], g[key + "_G"]) < 1e-12
    # sigma scaling and the drop-in surface
    trials = [dict(y=np.zeros((50, 3))), dict(y=np.zeros((120, 3)))]
    params["sigma"] = np.array([1.0, 0.5, 2.0])
    make_cholesky(trials, params, None)
    assert sorted(params["cholesky"]) == [50, 120] and params["cholesky"][120].shape == (3, 120, 50)
    for l in range(3):
        assert relerr(params["cholesky"][120][l], orc.ichol_gauss(120, omegas[l], 50) * params["sigma"][l]) < 1e-12


def test_ichol_known_answer_full_rank(vl):
    """The reference's own known-answer test (tests/test_math.py:7-14) at rank 60 <= VLGP_MAX_RANK: G G' == K."""
    eng = __import__("vlgp_b200.engine", fromlist=["get_engine"]).get_engine()
    n, omega = 60, 1.0
    params = dict(ydim=2, zdim=1, xdim=1, rank=60, gp_noise=1e-4, dt=1, likelihood=np.array(["poisson"] * 2),
                  sigma=np.ones(1), omega=np.array([omega]))
    eng.ensure_model(params)
    eng.push_params(params, which=("sigma", "omega"))
    with eng.new_trials([n]) as ts:
        ts.make_cholesky()
        G = ts.get_cholesky(n)[0]
    K = np.exp(-omega * (np.arange(n)[:, None] - np.arange(n)[None, :]) ** 2)
    assert np.allclose(K, G @ G.T)
    assert relerr(G, load_golden("ichol")["fullrank_n60_G"]) < 1e-10


@pytest.mark.parametrize("case", ["poisson_it1", "poisson_it25", "mixed_it1", "mixed_it25", "map_it3"])
def test_estep_golden(vl, case):
    from vlgp_b200 import core

    g = load_golden("estep")
    p = case + "_"
    segs = _segs(g, p)
    params = _params(g, p)
    params["cholesky"] = {50: g[p + "G"]}
    cfg = _cfg(Eniter=int(case.split("it")[1]), method="MAP" if case.startswith("map") else "VB")
    core.estep(segs, params, cfg)
    for k in ("mu", "v", "w", "dmu"):
        got = np.stack([s[k] for s in segs])
        scale = g[p + "out_mu"] if k == "dmu" else g[p + "out_" + k]
        assert np.max(np.abs(got - g[p + "out_" + k])) <= STEP_TOL * np.max(np.abs(scale)), k


@pytest.mark.parametrize("case", ["poisson_it1", "poisson_it25", "mixed_it1", "mixed_it25"])
def test_mstep_golden(vl, case):
    from vlgp_b200 import core

    g = load_golden("mstep")
    p = case + "_"
    segs = _segs(g, p)
    params = _params(g, p)
    cfg = _cfg(Mniter=int(case.split("it")[1]))
    core.mstep(segs, params, cfg)
    for k in ("a", "b", "noise"):
        assert relerr(params[k], g[p + "out_" + k]) < STEP_TOL, k
    for k in ("da", "db"):      # the last Newton step shrinks towards 0 as the iterations converge: scale of a / b
        assert np.max(np.abs(params[k] - g[p + "out_" + k])) < STEP_TOL * np.max(np.abs(g[p + "out_" + k[1]])), k
    assert params["b"].shape == g[p + "out_b"].shape


def test_hstep_objective_golden(vl):
    from vlgp_b200.core import Session

    g = load_golden("hstep")
    n = g["mu"].shape[0]
    segs = [dict(y=np.zeros((50, 2)), mu=g["mu"][i].copy(), w=g["w"][i].copy(), v=np.zeros((50, 2))) for i in range(n)]
    params = dict(a=np.zeros((2, 2)), b=np.zeros((1, 2)), noise=np.ones(2), omega=g["omega0"].copy(), sigma=np.ones(2),
                  likelihood=np.array(["poisson"] * 2), zdim=2, ydim=2, xdim=1, rank=50, gp_noise=1e-4, dt=1)
    with Session(segs, params, upload_factors=False) as s:
        s.ts.hstep_prepare()
        for l in range(2):
            for i, om in enumerate(g["omegas"]):
                ll, dll, info = s.ts.hstep_objective(l, np.array([1.0, om, 1e-4]))
                assert info == 0
                assert abs(ll - g["ll_l%d" % l][i]) < 1e-9 * abs(g["ll_l%d" % l][i])
                assert abs(dll - g["dll_l%d" % l][i][1]) < 1e-7 * max(abs(g["dll_l%d" % l][i][1]), 1.0)


def test_hstep_whole_golden(vl):
    from vlgp_b200 import core

    g = load_golden("hstep")
    n = g["mu"].shape[0]
    segs = [dict(y=np.zeros((50, 2)), mu=g["mu"][i].copy(), w=g["w"][i].copy(), v=np.zeros((50, 2))) for i in range(n)]
    params = dict(a=np.zeros((2, 2)), b=np.zeros((1, 2)), noise=np.ones(2), omega=g["omega0"].copy(), sigma=np.ones(2),
                  likelihood=np.array(["poisson"] * 2), zdim=2, ydim=2, xdim=1, rank=50, gp_noise=1e-4, dt=1)
    cfg = _cfg()
    core.hstep(segs, params, cfg)
    assert relerr(params["omega"], g["omega_after"]) < 1e-6
    assert relerr(params["sigma"], g["sigma_after"]) < 1e-12
    Gd = params["cholesky"][50]
    assert relerr(Gd @ Gd.transpose(0, 2, 1), g["G_after"] @ g["G_after"].transpose(0, 2, 1)) < 1e-5


@pytest.mark.parametrize("W", [17, 18, 26, 33, 40, 42, 49, 50, 56, 57, 64, 72, 100, 131])
def test_hstep_objective_window_lengths_vs_oracle(vl, W):
    """The H-step objective over window lengths that select every per-segment kernel: the bordered DMMA kernel (one or
    two bins beyond a multiple of 8: 17, 18, 26, 33, 42, 49, 50), the plain DMMA kernel (40, 56), the register sweep
    (57, 64) and the shared-memory kernels of csrc/hstep_wide.cu for windows of more than 64 bins (the reference takes
    any window, vlgp/gp.py:65-123) -- against the oracle's literal restatement of gp.elbo / construct_posterior_cov."""
    from vlgp_b200.core import Session
    from oracle import vlgp_oracle as orc

    rng = np.random.default_rng(W)
    S, L = 5, 2
    t = np.arange(W, dtype=float)
    mu = np.stack([np.stack([np.sin(2 * np.pi * (l + 1) * t / 60 + rng.uniform(0, 6)) for l in range(L)], axis=1)
                   + 0.1 * rng.standard_normal((W, L)) for _ in range(S)])
    w = rng.uniform(0.05, 3.0, size=(S, W, L))
    segs = [dict(y=np.zeros((W, 2)), mu=mu[i].copy(), w=w[i].copy(), v=np.zeros((W, L))) for i in range(S)]
    params = dict(a=np.zeros((L, 2)), b=np.zeros((1, 2)), noise=np.ones(2), omega=np.full(L, 1e-2), sigma=np.ones(L),
                  likelihood=np.array(["poisson"] * 2), zdim=L, ydim=2, xdim=1, rank=50, gp_noise=1e-4, dt=1)
    with Session(segs, params, upload_factors=False) as s:
        s.ts.hstep_prepare()
        for l in range(L):
            for om in (2e-3, 1e-2, 4e-2):
                hyper = np.array([1.0, om, 1e-4])
                ll, dll, info = s.ts.hstep_objective(l, hyper)
                f, g = orc.hstep_objective(np.log(hyper), t, mu[:, :, l].T, w[:, :, l].T)
                assert info == 0
                assert abs(ll + f) < 1e-9 * abs(f), (W, l, om)
                assert abs(dll + g[1]) < 1e-7 * max(abs(g[1]), 1.0), (W, l, om)


def test_fit_wide_window_golden(vl, monkeypatch):
    """fit(..., window=100) -- refused before this round -- against the reference.  Free-running: omega itself at 1e-5
    (the optimiser's sensitivity, as in test_fit_tutorial_golden).  With the reference's omega of every iteration
    injected AND the oracle's prior factors (at these omega the device's incomplete Cholesky breaks an exact pivot tie
    the other way, which alone moves mu by 5e-3): everything at 1e-7."""
    from conftest import inject_hyperparameter_trajectory, use_oracle_prior_factors
    from vlgp_b200.synth import make_trials

    g = load_golden("fit_wide_window")
    trials = make_trials(4, 200, 12, 2, seed=5)
    assert np.array_equal(np.stack([t["y"] for t in trials]), g["y"])
    np.random.seed(0)
    res = vl.fit(trials, 2, window=100, max_iter=3, min_iter=3)
    assert relerr(res["params"]["omega"], g["omega"]) < 1e-5
    # At this problem's final omega the rank-truncated factor of the T = 200 trials has an exact tie between two
    # mirror-image pivots; which one wins depends on the last digit of omega and moves mu by 5.07e-3 (both outcomes have
    # been observed here, with M-step summation orders that differ only in rounding).  Hence the pivot-flip scale for
    # the free-running fit, and 1e-7 below once omega and the factors are the reference's.
    assert relerr(np.stack([t["mu"] for t in res["trials"]]), g["mu"]) < 2e-2
    state = inject_hyperparameter_trajectory(monkeypatch, g["omega_traj"], g["sigma_traj"])
    use_oracle_prior_factors(monkeypatch)
    trials = make_trials(4, 200, 12, 2, seed=5)
    np.random.seed(0)
    res = vl.fit(trials, 2, window=100, max_iter=3, min_iter=3)
    assert state["it"] == 3
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in res["trials"]]), g[k]) < 1e-7, k
    for k in ("a", "b", "noise"):
        assert relerr(res["params"][k], g[k]) < 1e-7, k


def test_hstep_native_optimizer_against_scipy_on_the_device_objective(vl, monkeypatch):
    """vlgp_hstep_optimize (all L-BFGS-B rounds inside one native call, csrc/hstep_opt.cu) against scipy's own setulb
    driven from Python (VLGP_HSTEP_SCIPY=1) on the SAME device objective: without the collapse rule the two must ask
    for the same number of evaluations and end at the same point; with it (default, 1e-9) the end point stays within
    1e-6 of scipy's (vlgp/gp.py:100-123) while the device rounds drop."""
    from vlgp_b200 import core

    g = load_golden("hstep")
    n = g["mu"].shape[0]

    def run(**cfg_kw):
        segs = [dict(y=np.zeros((50, 2)), mu=g["mu"][i].copy(), w=g["w"][i].copy(), v=np.zeros((50, 2)))
                for i in range(n)]
        params = dict(a=np.zeros((2, 2)), b=np.zeros((1, 2)), noise=np.ones(2), omega=g["omega0"].copy(),
                      sigma=np.ones(2), likelihood=np.array(["poisson"] * 2), zdim=2, ydim=2, xdim=1, rank=50,
                      gp_noise=1e-4, dt=1)
        cfg = _cfg()
        cfg.update(cfg_kw)
        core.hstep(segs, params, cfg)
        return params["omega"].copy(), cfg["hstep_nfev"][-1], cfg.get("hstep_rounds", [None])[-1]

    monkeypatch.setenv("VLGP_HSTEP_SCIPY", "1")
    om_scipy, nf_scipy, _ = run()
    monkeypatch.delenv("VLGP_HSTEP_SCIPY")
    om_full, nf_full, rounds_full = run(hstep_collapse_tol=0.0)
    om_def, nf_def, rounds_def = run()
    assert list(nf_full) == list(nf_scipy)
    assert relerr(np.log(om_full), np.log(om_scipy)) < 1e-12
    assert np.max(np.abs(np.log(om_def) - np.log(om_scipy))) < 1e-6
    assert rounds_def <= rounds_full <= max(nf_scipy)
    assert relerr(om_def, g["omega_after"]) < 1e-6


def test_update_w_v_and_long_trial_estep_golden(vl):
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky

    g = load_golden("update_wv")
    y = g["y"].astype(float)
    N = y.shape[2]
    trials = [dict(y=y[i], x=np.ones((y.shape[1], 1, N)), mu=g["in_mu"][i].copy(), v=g["in_v"][i].copy())
              for i in range(y.shape[0])]
    L = g["a"].shape[0]
    params = dict(a=g["a"].copy(), b=g["b"].copy(), noise=g["noise"].copy(), omega=g["omega"].copy(),
                  sigma=g["sigma"].copy(), zdim=L, ydim=N, xdim=1, rank=50, gp_noise=1e-4, dt=1,
                  likelihood=np.array(["poisson"] * N))
    cfg = _cfg(Eniter=3)
    make_cholesky(trials, params, cfg)
    core.update_w(trials, params, cfg)
    core.update_v(trials, params, cfg)
    assert relerr(np.stack([t["w"] for t in trials]), g["out_w"]) < STEP_TOL
    assert relerr(np.stack([t["v"] for t in trials]), g["out_v"]) < STEP_TOL
    for t in trials:
        t["dmu"] = np.zeros_like(t["mu"])
    core.estep(trials, params, cfg)
    for k in ("mu", "v", "w", "dmu"):
        scale = g["infer_mu"] if k == "dmu" else g["infer_" + k]
        assert np.max(np.abs(np.stack([t[k] for t in trials]) - g["infer_" + k])) < STEP_TOL * np.max(np.abs(scale)), k


@pytest.mark.parametrize("rank,lengths", [(30, [300, 77, 513]), (50, [64, 256, 257, 1000]), (8, [120, 120])])
def test_long_trial_estep_vs_oracle(vl, rank, lengths):
    """Long-trial E-step kernels (csrc/estep_long.cu) against the oracle: item boundaries (lengths that are / are not
    multiples of the 64-bin chunks and 256-bin items), the 32-column instantiation (rank <= 32) next to the 56-column one,
    several trials per length, MAP as well as VB."""
    from vlgp_b200 import core
    from oracle import vlgp_oracle as orc

    rng = np.random.default_rng(rank)
    N, L = 40, 4
    params = dict(a=0.15 * rng.standard_normal((L, N)), b=np.full((1, N), np.log(0.1)), noise=np.ones(N),
                  omega=np.exp(rng.uniform(np.log(2e-3), np.log(3e-2), L)), sigma=np.ones(L),
                  likelihood=np.array(["poisson"] * N), zdim=L, ydim=N, xdim=1, rank=rank, gp_noise=1e-4, dt=1)
    trials = []
    for n in lengths:
        trials.append(dict(y=rng.poisson(0.15, (n, N)).astype(float), x=np.ones((n, 1, N)),
                           mu=0.2 * rng.standard_normal((n, L)), v=np.zeros((n, L)), w=np.zeros((n, L)),
                           dmu=np.zeros((n, L))))
    params["cholesky"] = orc.make_cholesky(sorted(set(lengths)), params["omega"], params["sigma"], rank)
    for method in ("VB", "MAP"):
        cfg = _cfg(Eniter=4, method=method)
        t_ref, t_dev = copy.deepcopy(trials), copy.deepcopy(trials)
        orc.update_w(t_ref, params)
        orc.update_v(t_ref, params, cfg)
        for a, b in zip(t_dev, t_ref):
            a["w"], a["v"] = b["w"].copy(), b["v"].copy()
        orc.estep(t_ref, copy.deepcopy(params), cfg)
        core.estep(t_dev, copy.deepcopy(params), cfg)
        for a, b in zip(t_dev, t_ref):
            for k in ("mu", "v", "w", "dmu"):
                scale = b["mu"] if k == "dmu" else b[k]
                assert np.max(np.abs(a[k] - b[k])) <= 1e-9 * max(np.max(np.abs(scale)), 1e-300), (method, k, a[k].shape)


def test_vem_three_iterations_golden(vl):
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky

    g = load_golden("vem")
    segs = _segs(g, "")
    N = g["y"].shape[2]
    params = _params(g, "", poisson=np.ones(N, bool))
    cfg = _cfg(max_iter=3, min_iter=3)
    make_cholesky(segs, params, cfg)
    core.vem(segs, params, cfg)
    assert cfg["runtime"]["it"] == int(g["n_it"])
    assert len(cfg["runtime"]["em_elapsed"]) == 3
    assert relerr(params["omega"], g["out_omega"]) < 1e-5
    for k in ("a", "b"):
        assert relerr(params[k], g["out_" + k]) < 1e-5, k
    assert relerr(np.stack([s["mu"] for s in segs]), g["out_mu"]) < 1e-5
    assert relerr(np.stack([s["v"] for s in segs]), g["out_v"]) < 1e-5


def test_fit_tutorial_golden(vl):
    """Whole fit() on BASELINE config 1 (10 x 200 x 30 x 3) against the reference run with the same global seed."""
    from vlgp_b200.synth import make_trials

    g = load_golden("fit_tutorial")
    trials = make_trials(10, 200, 30, 3, seed=0)
    assert np.array_equal(np.stack([t["y"] for t in trials]), g["y"])
    np.random.seed(0)
    res = vl.fit(trials, 3, max_iter=3, min_iter=3)
    assert set(res) == {"trials", "params", "config"}
    assert relerr(res["params"]["initial"]["a"], g["initial_a"]) < 1e-9      # same FactorAnalysis initialisation
    # With the H-step on, omega comes out of L-BFGS-B, whose iterates move by ~1e-11 when the objective moves by 1 ulp
    # (measured: identical inputs to 1e-16 give nfev [24,31,23] vs [34,21,20], scripts/debug_fit_divergence.py).  A
    # 1e-11 change of omega can flip one of the EXACT ties between mirror-image pivots of the incomplete Cholesky
    # (residual diagonal is symmetric in t <-> W-1-t), which swaps one column of the rank-truncated prior factor: an
    # O(tol = 1e-6..1e-4) change of the model that no implementation can avoid (hard parts 1 and
    # 6).  The tight end-to-end check therefore runs with omega fixed (test_fit_fixed_omega_golden, 1e-7); here the
    # bound is the pivot-flip scale.
    mu = np.stack([t["mu"] for t in res["trials"]])
    assert relerr(mu, g["mu"]) < 5e-4
    assert relerr(res["params"]["a"], g["a"]) < 5e-4
    assert relerr(res["params"]["b"], g["b"]) < 5e-4
    assert relerr(res["params"]["omega"], g["omega"]) < 1e-5
    assert relerr(np.stack([t["v"] for t in res["trials"]]), g["v"]) < 5e-4
    for k in ("a", "b", "noise", "sigma", "omega", "da", "db", "cholesky", "rank", "gp_noise", "dt", "likelihood",
              "xdim", "ydim", "zdim", "transform", "initial"):
        assert k in res["params"], k
    for k in ("mu", "v", "w", "dmu", "x", "cut", "y", "ID"):
        assert k in res["trials"][0], k
    assert res["params"]["cholesky"][200].shape == (3, 200, 50)
    assert set(res["config"]["runtime"]) >= {"it", "e_elapsed", "m_elapsed", "h_elapsed", "em_elapsed"}


def test_default_fit_with_the_reference_omega_trajectory(vl, monkeypatch):
    """North-star tolerance on the DEFAULT path: fit() with the H-step ON, where the optimiser's result of each EM
    iteration is replaced by the reference's own (tests/golden/fit_tutorial.npz, omega_traj).  Prior factors are rebuilt
    on the device from each new omega (ichol_gauss, vlgp/math.py:76-126), the E- / M-steps and the final full-trial
    inference run on them: posterior means within 1e-7 of the reference (north star: 1e-5).  What this leaves out is
    only L-BFGS-B's amplification of last-digit differences of its objective (test_fit_tutorial_golden, 5e-4)."""
    from conftest import inject_hyperparameter_trajectory
    from vlgp_b200.synth import make_trials

    g = load_golden("fit_tutorial")
    state = inject_hyperparameter_trajectory(monkeypatch, g["omega_traj"], g["sigma_traj"])
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = vl.fit(trials, 3, max_iter=3, min_iter=3)
    assert state["it"] == 3
    assert np.array_equal(res["params"]["omega"], g["omega"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in res["trials"]]), g[k]) < 1e-7, k
    for k in ("a", "b", "noise"):
        assert relerr(res["params"][k], g[k]) < 1e-7, k


def test_fit_fixed_omega_golden(vl):
    """Whole fit() (FactorAnalysis init, update_w/v, cut, 3 EM iterations of E+M, final infer on the uncut trials) with
    the H-step off, against the reference run with the same global seed: posterior means within 1e-7 relative (the
    north-star tolerance is 1e-5)."""
    from vlgp_b200.synth import make_trials

    g = load_golden("fit_fixed_omega")
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = vl.fit(trials, 3, max_iter=3, min_iter=3, Hstep=False)
    assert np.array_equal(res["params"]["omega"], g["omega"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in res["trials"]]), g[k]) < 1e-7, k
    for k in ("a", "b", "noise"):
        assert relerr(res["params"][k], g[k]) < 1e-7, k


# ----------------------------------------------------------------------------------------------------------------------
# seeded comparisons with the oracle at sizes it finishes in seconds, and properties at full size
# ----------------------------------------------------------------------------------------------------------------------
def _problem(seed, n_trials, T, N, L, lik=None, window=50, a_scale=0.3):
    from vlgp_b200.synth import make_trials
    from oracle import vlgp_oracle as orc

    rng = np.random.default_rng(seed)
    trials = make_trials(n_trials, T, N, L, seed=seed + 100)
    poisson = np.ones(N, bool) if lik is None else np.asarray(lik) == "poisson"
    params = dict(a=a_scale * rng.standard_normal((L, N)), b=np.full((1, N), np.log(0.08)), noise=0.5 + rng.random(N),
                  omega=np.exp(rng.uniform(np.log(2e-3), np.log(4e-2), L)), sigma=np.ones(L),
                  likelihood=np.where(poisson, "poisson", "gaussian"), zdim=L, ydim=N, xdim=1, rank=50, gp_noise=1e-4,
                  dt=1)
    params["da"] = np.zeros_like(params["a"])
    params["db"] = np.zeros_like(params["b"])
    segs = []
    for tr in trials:
        for s in range(0, tr["y"].shape[0] - window + 1, window):
            y = tr["y"][s:s + window].copy()
            y[:, ~poisson] += 0.3 * rng.standard_normal((window, int((~poisson).sum())))
            segs.append(dict(y=y, x=np.ones((window, 1, N)), mu=0.5 * rng.standard_normal((window, L)),
                             v=np.zeros((window, L)), w=np.zeros((window, L)), dmu=np.zeros((window, L))))
    params["cholesky"] = orc.make_cholesky([window], params["omega"], params["sigma"], 50)
    orc.update_w(segs, params)
    orc.update_v(segs, params, orc.default_config())
    return segs, params


@pytest.mark.parametrize("N,L,lik", [(100, 5, None), (37, 10, None), (20, 4, ["poisson"] * 12 + ["gaussian"] * 8)])
def test_estep_mstep_vs_oracle_seeded(vl, N, L, lik):
    from vlgp_b200 import core
    from oracle import vlgp_oracle as orc

    segs, params = _problem(11, 2, 150, N, L, lik)
    cfg = _cfg(Eniter=4, Mniter=3)
    s_ref, p_ref = copy.deepcopy(segs), copy.deepcopy(params)
    orc.estep(s_ref, p_ref, cfg)
    # conditioning of this instance: how far the ORACLE's own output moves when its input mu moves by 1e-15 relative
    # (random far-from-converged states with clipped Newton steps can amplify rounding by ~1e7; two correct fp64
    # implementations cannot agree better than that)
    s_pert = copy.deepcopy(segs)
    for sg in s_pert:
        sg["mu"] = sg["mu"] * (1 + 1e-15)
    orc.estep(s_pert, copy.deepcopy(params), cfg)
    mu_ref = np.stack([s["mu"] for s in s_ref])
    sens = np.max(np.abs(np.stack([s["mu"] for s in s_pert]) - mu_ref)) / np.max(np.abs(mu_ref))
    tol = max(STEP_TOL, 10 * sens)
    core.estep(segs, params, cfg)
    for k in ("mu", "v", "w", "dmu"):
        ref = np.stack([s[k] for s in s_ref])
        scale = mu_ref if k == "dmu" else ref
        assert np.max(np.abs(np.stack([s[k] for s in segs]) - ref)) <= tol * np.max(np.abs(scale)), (k, tol)
    for sg, sr in zip(segs, s_ref):         # continue both M-steps from identical inputs
        for k in ("mu", "v", "w", "dmu"):
            sg[k] = sr[k].copy()
    orc.mstep(s_ref, p_ref, cfg)
    core.mstep(segs, params, cfg)
    for k in ("a", "b", "noise"):
        assert relerr(params[k], p_ref[k]) < STEP_TOL, k
    for k in ("da", "db"):
        assert np.max(np.abs(params[k] - p_ref[k])) < STEP_TOL * np.max(np.abs(p_ref[k[1]])), k


@pytest.mark.parametrize("N,L,T,window", [(100, 5, 45, 15), (30, 3, 150, 50), (200, 10, 75, 25), (600, 2, 63, 21),
                                          (7, 1, 33, 11)])
def test_mstep_tma_staged_statistics_vs_oracle(vl, N, L, T, window):
    """The TMA-staged statistics kernel of the middle Newton iterations (csrc/mstep.cu) on awkward geometries: odd
    bin counts (a single bin left over at the end of the range), ranges shorter than one stage, one and many bin
    lanes per CTA, more neurons than one CTA holds.  Reference: vlgp/core.py:129-249."""
    from vlgp_b200 import core
    from oracle import vlgp_oracle as orc

    segs, params = _problem(23, 3, T, N, L, None, window=window)
    cfg = _cfg(Eniter=2, Mniter=6)
    s_ref, p_ref = copy.deepcopy(segs), copy.deepcopy(params)
    orc.mstep(s_ref, p_ref, cfg)
    core.mstep(segs, params, cfg)
    for k in ("a", "b", "noise"):
        assert relerr(params[k], p_ref[k]) < STEP_TOL, k
    for k in ("da", "db"):
        assert np.max(np.abs(params[k] - p_ref[k])) < STEP_TOL * np.max(np.abs(p_ref[k[1]])), k


def test_estep_config2_shape_subset_vs_oracle(vl):
    """BASELINE config 2 shape (T=1000 trials cut in 50-bin windows, N=100, L=5): all 5120 segments run on the GPU; a
    random subset of 6 segments is checked against the oracle, and every segment against invariants."""
    from vlgp_b200 import core
    from oracle import vlgp_oracle as orc

    segs, params = _problem(5, 256, 1000, 100, 5)
    assert len(segs) == 5120
    cfg = _cfg(Eniter=25)
    pick = np.random.default_rng(0).choice(len(segs), 6, replace=False)
    s_ref = [copy.deepcopy(segs[i]) for i in pick]
    orc.estep(s_ref, copy.deepcopy(params), cfg)
    core.estep(segs, params, cfg)
    for j, i in enumerate(pick):
        for k in ("mu", "v", "w"):
            assert relerr(segs[i][k], s_ref[j][k]) < STEP_TOL, (i, k)
    v = np.stack([s["v"] for s in segs])
    w = np.stack([s["w"] for s in segs])
    prior_var = np.einsum("ltr,ltr->tl", params["cholesky"][50], params["cholesky"][50])
    assert np.all(v > 0) and np.all(v <= prior_var[None] * (1 + 1e-12))     # posterior variance below the prior's
    assert np.all(w > 0) and np.all(np.isfinite(np.stack([s["mu"] for s in segs])))
    # update_w on all 5120 segments reproduces the oracle's weights on the subset (w then includes the new v)
    core.update_w(segs, params, cfg)
    orc.update_w(s_ref, params)
    for j, i in enumerate(pick):
        assert relerr(segs[i]["w"], s_ref[j]["w"]) < STEP_TOL


F32_TOL = 1e-5          # posterior means, single-precision rate passes (north_star: "within 1e-5 relative of reference")


@pytest.mark.parametrize("shape", ["golden_poisson_it25", "config2", "config3"])
def test_estep_float32_rate_passes(vl, shape):
    """config["dtype"] = "float32" (BASELINE.json configs[2]): the segment E-step with its rate passes in single precision
    (csrc/estep_seg_impl.cuh: rate_tiles_f32; Gram / inverse / variance / mean step stay double).  Posterior means within
    F32_TOL of the reference's output (golden from vlgp/core.py:22-120) and of the oracle on the config-2 and config-3
    shapes (N = 200, L = 10); variances and weights within 1e-4.  The result must DIFFER from the double-precision run
    (the switch did something) and dtype="float64" afterwards must be bit-identical to a run that never switched."""
    from vlgp_b200 import core
    from oracle import vlgp_oracle as orc

    if shape.startswith("golden"):
        g = load_golden("estep")
        p = "poisson_it25_"
        segs, params = _segs(g, p), _params(g, p)
        params["cholesky"] = {50: g[p + "G"]}
        ref = {k: g[p + "out_" + k] for k in ("mu", "v", "w")}
    else:
        # (at L = 10 a loading of scale 0.3 makes the Jacobi-over-latents Newton iteration of some segments oscillate
        # against the +-dmu_bound clip: rounding differences between any two implementations are then amplified to
        # 1e-2 within 25 iterations, double precision included; scale 0.1 is a contracting iteration)
        N, L, sc = (100, 5, 0.3) if shape == "config2" else (200, 10, 0.1)
        segs, params = _problem(11, 4, 1000, N, L, a_scale=sc)
        s_ref = copy.deepcopy(segs)
        orc.estep(s_ref, copy.deepcopy(params), _cfg(Eniter=25))
        ref = {k: np.stack([s[k] for s in s_ref]) for k in ("mu", "v", "w")}
    s64, s32, s64b = copy.deepcopy(segs), copy.deepcopy(segs), copy.deepcopy(segs)
    core.estep(s64, copy.deepcopy(params), _cfg(Eniter=25))
    core.estep(s32, copy.deepcopy(params), _cfg(Eniter=25, dtype="float32"))
    core.estep(s64b, copy.deepcopy(params), _cfg(Eniter=25))
    got32 = {k: np.stack([s[k] for s in s32]) for k in ("mu", "v", "w")}
    got64 = {k: np.stack([s[k] for s in s64]) for k in ("mu", "v", "w")}
    assert relerr(got64["mu"], ref["mu"]) < STEP_TOL
    err = {k: relerr(got32[k], ref[k]) for k in ref}
    assert err["mu"] < F32_TOL, err
    assert err["v"] < 1e-4 and err["w"] < 1e-4, err
    assert relerr(got32["mu"], got64["mu"]) > 1e-12                      # single precision really ran
    for a, b in zip(s64, s64b):
        assert np.array_equal(a["mu"], b["mu"])                          # and did not leak into the default path


def test_overlapping_windows_and_unequal_lengths(vl):
    """T not a multiple of the window (config 3 regime): segments overlap, state is written back through views, and
    the final infer runs on trials of different lengths (one prior factor per unique length)."""
    from vlgp_b200.synth import make_trials

    trials = make_trials(3, (130, 170), 12, 2, seed=3)
    lengths = [t["y"].shape[0] for t in trials]
    assert len(set(lengths)) > 1
    np.random.seed(1)
    res = vl.fit(trials, 2, max_iter=2, min_iter=2)
    assert sorted(res["params"]["cholesky"]) == sorted(set(lengths))
    for t in res["trials"]:
        assert t["mu"].shape == (t["y"].shape[0], 2) and np.all(np.isfinite(t["mu"])) and np.all(t["v"] > 0)


def test_config3_shape_long_unequal_trials_vs_oracle(vl):
    """BASELINE config 3 shape at reduced trial count: N = 200 neurons, L = 10 latents, unequal lengths in [500, 2000]
    (fp64 here; the fp32 arithmetic of config 3 is not built).  The uncut-trial path (one prior factor per unique
    length, T x 50 factors streamed from HBM) is compared with the oracle on every trial."""
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky
    from vlgp_b200.synth import make_trials
    from oracle import vlgp_oracle as orc

    rng = np.random.default_rng(3)
    N, L = 200, 10
    trials = make_trials(3, (500, 2000), N, L, seed=7)
    lengths = [t["y"].shape[0] for t in trials]
    assert len(set(lengths)) == 3
    params = dict(a=0.1 * rng.standard_normal((L, N)), b=np.full((1, N), np.log(0.08)), noise=np.ones(N),
                  omega=np.exp(rng.uniform(np.log(2e-3), np.log(4e-2), L)), sigma=np.ones(L),
                  likelihood=np.array(["poisson"] * N), zdim=L, ydim=N, xdim=1, rank=50, gp_noise=1e-4, dt=1)
    for t in trials:
        n = t["y"].shape[0]
        t.update(x=np.ones((n, 1, N)), mu=0.1 * rng.standard_normal((n, L)), v=np.zeros((n, L)), w=np.zeros((n, L)),
                 dmu=np.zeros((n, L)))
    cfg = _cfg(Eniter=3)
    make_cholesky(trials, params, cfg)
    ref_chol = orc.make_cholesky(lengths, params["omega"], params["sigma"], 50)
    # The device factor equals the oracle's whenever the pivot sequences coincide.  They can differ where two
    # candidate pivots tie to the last ulp (about 1 factorisation in 10 at random omega, scripts/debug_ichol_pivots.py):
    # NumPy/OpenBLAS and the GPU round the recurrences differently, so the "first arg-max" falls on another row.  Both
    # factors are then equally good rank-50 approximations of K (checked below); the E-step comparison that follows
    # injects the oracle's factors so that it does not depend on which tie-break happened.
    n_same = 0
    for T in lengths:
        K = np.exp(-params["omega"][:, None, None] * (np.arange(T)[None, :, None] - np.arange(T)[None, None, :]) ** 2.0)
        for l in range(L):
            Gd, Gr = params["cholesky"][T][l], ref_chol[T][l]
            if relerr(Gd, Gr) < 1e-11:
                n_same += 1
            else:
                ed, er = np.abs(Gd @ Gd.T - K[l]).max(), np.abs(Gr @ Gr.T - K[l]).max()
                assert ed < 1.5 * er + 1e-9, (T, l, ed, er)
    assert n_same >= 0.6 * len(lengths) * L
    params["cholesky"] = {T: ref_chol[T].copy() for T in lengths}
    t_ref, p_ref = copy.deepcopy(trials), copy.deepcopy(params)
    orc.update_w(t_ref, p_ref)
    orc.update_v(t_ref, p_ref, cfg)
    orc.estep(t_ref, p_ref, cfg)
    core.update_w(trials, params, cfg)
    core.update_v(trials, params, cfg)
    core.estep(trials, params, cfg)
    for a, b in zip(trials, t_ref):
        for k in ("mu", "v", "w"):
            assert relerr(a[k], b[k]) < STEP_TOL, k


def test_callbacks_and_constraints(vl):
    """vem surface details: callbacks see coherent host dicts every iteration; constrain_latent='both' and the 'svd' /
    row-norm loading constraints run on the device and match the oracle."""
    from vlgp_b200 import core
    from oracle import vlgp_oracle as orc

    for loading, latent in (("fro", "both"), ("svd", False), (2, "location")):
        segs, params = _problem(21, 2, 100, 15, 3)
        cfg = _cfg(max_iter=2, min_iter=2, Eniter=3, Mniter=3, Hstep=False, constrain_loading=loading,
                   constrain_latent=latent)
        s_ref, p_ref, c_ref = copy.deepcopy(segs), copy.deepcopy(params), copy.deepcopy(cfg)
        seen = []
        cfg["callbacks"] = [lambda tr, pa, co: seen.append((co["runtime"]["it"], float(np.sum(tr[0]["mu"]))))]
        core.vem(segs, params, cfg)
        orc.vem(s_ref, p_ref, c_ref)
        assert [it for it, _ in seen] == [1, 2] and np.isfinite(seen[-1][1])
        assert relerr(np.stack([s["mu"] for s in segs]), np.stack([s["mu"] for s in s_ref])) < 1e-8, (loading, latent)
        assert relerr(params["a"], p_ref["a"]) < 1e-8 and relerr(params["b"], p_ref["b"]) < 1e-8


@pytest.mark.parametrize("N,L,window,big_counts", [(3, 1, 50, False), (17, 2, 30, False), (12, 3, 64, False),
                                                     (9, 2, 50, True), (40, 7, 24, False)])
def test_vem_shapes_vs_oracle(vl, N, L, window, big_counts):
    """Whole EM iterations (constrain + E + M + H) at odd shapes: one latent, windows that are not a multiple of the
    8 x 8 tensor tile, the largest supported window, counts above 255 (float64 storage), seven latents."""
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky
    from oracle import vlgp_oracle as orc

    segs, params = _problem(31 + N, 2, 4 * window, N, L, window=window)
    if big_counts:
        segs[0]["y"][3, 1] = 300.0
    cfg = _cfg(max_iter=2, min_iter=2, Eniter=4, Mniter=3, window=window)
    params["cholesky"] = orc.make_cholesky([window], params["omega"], params["sigma"], 50)
    for sgm in segs:
        sgm["v"][...] = 0.0
        sgm["w"][...] = 0.0
    orc.update_w(segs, params)
    orc.update_v(segs, params, cfg)
    s_ref, p_ref, c_ref = copy.deepcopy(segs), copy.deepcopy(params), copy.deepcopy(cfg)
    core.vem(segs, params, cfg)
    orc.vem(s_ref, p_ref, c_ref)
    # omega is the optimum of a flat objective (L-BFGS-B stops on a 2e-9 relative decrease): 1e-4 over two iterations
    assert relerr(params["omega"], p_ref["omega"]) < 1e-4
    # the oracle and the device may break exact pivot ties of the new prior factor differently:
    # compare what the factor is used for, G G', and the posterior to the truncation level of the factor
    Gd, Gr = params["cholesky"][window], p_ref["cholesky"][window]
    same_factor = relerr(Gd, Gr) < 1e-9
    tol = 1e-7 if same_factor else 5e-4
    assert relerr(params["a"], p_ref["a"]) < tol and relerr(params["b"], p_ref["b"]) < tol
    assert relerr(np.stack([s["mu"] for s in segs]), np.stack([s["mu"] for s in s_ref])) < tol
    assert relerr(np.stack([s["v"] for s in segs]), np.stack([s["v"] for s in s_ref])) < tol


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
def test_vem_option_branches_golden(vl, case):
    """Two vem iterations against the reference's outputs (tests/golden/vem_options.npz) under the options the default
    fit never takes; the CPU suite pins the oracle to the same vectors (tests/test_oracle_golden.py)."""
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky

    g = load_golden("vem_options")
    p = case + "/"
    segs = _segs(g, p)
    params = _params(g, p)
    cfg = _cfg(max_iter=2, min_iter=2, **VEM_OPTION_CASES[case])
    make_cholesky(segs, params, cfg)
    core.vem(segs, params, cfg)
    assert cfg["runtime"]["it"] == int(g[p + "n_it"])
    tol = 5e-4 if cfg["Hstep"] else 1e-8      # with the H-step: L-BFGS-B end point + pivot ties of the new factor
    if cfg["method"] == "MAP":
        # without the variance term the Newton iteration on the posterior mean amplifies rounding differences by about
        # 1e3 per EM iteration on this kind of data (CPU against CPU: reference vs NumPy port 4e-14 / 5e-11 / 6e-8 after
        # 1 / 2 / 3 iterations, scripts/fuzz_options_vs_reference.py), so last-digit differences of the device arithmetic
        # do not stay at 1e-13
        tol = max(tol, 1e-6)
    assert relerr(params["omega"], g[p + "out_omega"]) < 1e-4
    for k in ("a", "b"):
        assert relerr(params[k], g[p + "out_" + k]) < tol, k
    for k in ("mu", "v"):
        assert relerr(np.stack([s[k] for s in segs]), g[p + "out_" + k]) < tol, k


def test_transform_new_trials_golden(vl):
    """fit(Hstep=False) then transform() of three trials the model has not seen, against the reference
    (tests/golden/api_extras.npz).  transform() starts from FactorAnalysis.transform evaluated with the FITTED loading
    (params['a'] is the estimator's components_ array and is updated in place, see tests/test_host_alias.py)."""
    import vlgp_b200 as vlgp
    from vlgp_b200.synth import make_trials

    g = load_golden("api_extras")
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = vlgp.fit(trials, 3, max_iter=3, min_iter=3, Hstep=False)
    new = make_trials(3, 200, 30, 3, seed=77)
    vlgp.transform(new, res["params"], res["config"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in new]), g["new_" + k]) < 1e-6, k


VEM_REGRESSOR_CASES = {
    "history2_poisson": (3, 0.0, dict(Hstep=False)),
    "history1_mixed_hstep": (2, 0.0, dict()),
    "scaled_bias_only": (1, 0.5, dict(Hstep=False, use_hessian=False, learning_rate=1e-4)),
}


@pytest.mark.parametrize("case", sorted(VEM_REGRESSOR_CASES))
def test_vem_general_regressors_golden(vl, case):
    """Regressors other than the all-ones bias column (xdim = max(history, 1) > 1 or a user design; vlgp/core.py:66,
    205-220,229-235) against the reference's outputs (tests/golden/vem_regressors.npz): E-step with the offsets
    einsum(x, b), M-step Newton / least-squares update of b with the design x (csrc/regress.cu)."""
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky
    from test_oracle_golden import regressor_design

    g = load_golden("vem_regressors")
    p = case + "/"
    xdim, scale, kw = VEM_REGRESSOR_CASES[case]
    segs = _segs(g, p)
    for sg in segs:
        sg["x"] = regressor_design(sg["y"], xdim, scale)
    params = _params(g, p)
    params["xdim"] = xdim
    cfg = _cfg(max_iter=2, min_iter=2, **kw)
    make_cholesky(segs, params, cfg)
    core.vem(segs, params, cfg)
    assert params["b"].shape == (xdim, 10) and params["db"].shape == (xdim, 10)
    tol = 5e-4 if cfg["Hstep"] else 1e-8
    for k in ("a", "b", "noise"):
        assert relerr(params[k], g[p + "out_" + k]) < tol, k
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([s[k] for s in segs]), g[p + "out_" + k]) < tol, k


def test_posterior_cov_and_sample_posterior(vl):
    """api.posterior_cov / sample_posterior (vlgp/api.py:142-168, vlgp/util.py:541-547): the T x T covariance from the
    device against the reference's expression evaluated in NumPy, and draws with the right moments."""
    from oracle import vlgp_oracle as orc

    g = load_golden("api_extras")
    G = orc.make_cholesky([200], g["omega"], g["sigma"], 50)[200]
    params = {"cholesky": {200: G}, "gp_noise": 1e-4, "dt": 1}
    trial = {"mu": g["trial0_mu"], "w": g["trial0_w"]}
    for reg in (1e-6, 1e-3, 0.0):
        cov = vl.posterior_cov(trial, params, reg)
        assert cov.shape == (3, 200, 200)
        for k in range(3):
            K = G[k] @ G[k].T
            if reg > 0:
                ref = np.linalg.inv(np.linalg.inv(K + reg * np.eye(200)) + np.diag(trial["w"][:, k]))
            else:
                ref = K - K @ np.linalg.solve(np.diag(1.0 / trial["w"][:, k]) + K, K)      # util.posterior_cov
            assert relerr(cov[k], ref) < 5e-9, (reg, k)
            assert np.array_equal(cov[k], cov[k].T) or relerr(cov[k], cov[k].T) < 1e-13
    np.random.seed(5)
    samples = vl.sample_posterior(trial, params, 400)
    assert samples.shape == (400, 200, 3) and np.isfinite(samples).all()
    cov = vl.posterior_cov(trial, params, 1e-6)
    for k in range(3):
        sd = np.sqrt(np.diag(cov[k]))
        z = (samples[:, :, k].mean(axis=0) - trial["mu"][:, k]) / (sd / np.sqrt(400))
        assert np.abs(z).max() < 6.0                                   # sample means within 6 standard errors
        ratio = samples[:, :, k].var(axis=0) / np.diag(cov[k])
        assert 0.6 < np.median(ratio) < 1.4


@pytest.mark.parametrize("case", ["eye", "noise", "one"])
def test_gpfa_em_golden(vl, case):
    """GPFA branch: gpfa.em on the device (csrc/gpfa.cu) against the reference's em() (vlgp/gpfa.py:20-56), including a
    non-uniform initial R, which exposes the reference's (time, neuron) ordering of bigR."""
    from vlgp_b200 import gpfa

    g = load_golden("gpfa")
    p = case + "_"
    z, C, d, R = gpfa.em(g[p + "y"], g[p + "C"], g[p + "d"], g[p + "R"], g[p + "K"], int(g[p + "iters"]))
    for got, key in ((z, "z"), (C, "C"), (d, "d"), (R, "R")):
        assert relerr(got, g[p + "out_" + key]) < 1e-9, key
    assert relerr(gpfa.sekernel(np.arange(g[p + "K"].shape[0]) * 1.0, 1.0, 3.0), g[p + "K"]) < 1e-15


def test_gpfa_fit_and_fastfit_config2_shape(vl):
    """gpfa.fit on a config-2-shaped problem (window 50, 100 neurons, 5 latents) against the oracle's em on the same
    prepared inputs; infer() against the E-step it restates; api.fastfit runs end to end and leaves a posterior."""
    import vlgp_b200
    from vlgp_b200 import gpfa
    from vlgp_b200.synth import make_trials
    from oracle import vlgp_oracle as orc

    trials = make_trials(8, 200, 100, 5, seed=4)
    np.random.seed(0)
    y, C0, d0, R0, K = gpfa.prepare(copy.deepcopy(trials), 5, dt=1.0, var=1.0, scale=8.0)
    assert y.shape == (32, 50, 100) and K.shape == (50, 50)
    zr, Cr, dr, Rr = orc.gpfa_em(y, C0, d0, R0, K, 3)
    z, C, d, R = gpfa.em(y, C0, d0, R0, K, 3)
    for got, ref in ((z, zr), (C, Cr), (d, dr), (R, Rr)):
        assert relerr(got, ref) < 1e-9
    np.random.seed(0)
    y2, z2, C2, d2, R2 = gpfa.fit(copy.deepcopy(trials), 5, dt=1.0, var=1.0, scale=8.0, max_iter=3)
    assert relerr(z2, zr) < 1e-9 and relerr(C2, Cr) < 1e-9
    # infer: one short trial, against the dense expression of vlgp/gpfa.py:59-76
    tr = dict(y=trials[0]["y"][:30].astype(float), mu=np.zeros((30, 5)), K=gpfa.sekernel(np.arange(30) * 1.0, 1.0, 8.0))
    Rn = np.diag(0.5 + np.random.default_rng(1).random(100))
    gpfa.infer([tr], C, np.abs(d) + 0.1, Rn)
    n, ydim, zdim = 30, 100, 5
    bigC, bigK, bigR = np.kron(C.T, np.eye(n)), np.kron(np.eye(zdim), tr["K"]), np.kron(np.eye(n), Rn)
    A = bigK @ bigC.T
    zz = A @ np.linalg.solve(bigC @ A + bigR, (tr["y"] - (np.abs(d) + 0.1)).T.reshape(-1, 1))
    assert relerr(tr["mu"], zz.reshape((zdim, -1)).T) < 1e-9
    # fastfit: GPFA then vLGP inference from it
    np.random.seed(0)
    tf = copy.deepcopy(trials)
    assert vlgp_b200.api.fastfit(tf, 5, dt=1.0, var=1.0, scale=8.0, max_iter=2) is None
    assert all(np.isfinite(t["mu"]).all() and t["mu"].shape == (200, 5) and (t["v"] > 0).all() for t in tf)


def test_reference_api_smoke(vl):
    """The reference's own API test (tests/test_api.py:4-38 there) with `import vlgp_b200 as vlgp`: integer counts from
    np.random.poisson, an extra user key per trial, fit with every default, then transform on the fitted trials."""
    import vlgp_b200 as vlgp

    np.random.seed(3)
    ydim, zdim, length, ntrial = 5, 2, 100, 5
    a = np.random.randn(zdim, ydim)
    trials = []
    for i in range(ntrial):
        z = np.column_stack((np.sin(np.linspace(0, 8 * np.pi, length)), np.cos(np.linspace(0, 8 * np.pi, length))))
        trials.append({"y": np.random.poisson(np.exp(z @ a - 2)), "id": i})
    result = vlgp.fit(trials, n_factors=2)
    out, params, config = result["trials"], result["params"], result["config"]
    assert out is trials and [t["id"] for t in out] == list(range(ntrial))
    assert 5 <= config["runtime"]["it"] <= config["max_iter"]
    for t in out:
        assert t["y"].dtype.kind == "i"                      # the caller's array is not converted in place
        for k in ("mu", "v", "w", "dmu"):
            assert t[k].shape == (length, zdim) and np.isfinite(t[k]).all(), k
        assert (t["v"] > 0).all() and (t["w"] > 0).all()
    assert params["a"].shape == (zdim, ydim) and params["b"].shape == (1, ydim)
    assert set(params["cholesky"]) == {length}
    mu_fit = np.stack([t["mu"] for t in out])
    vlgp.transform(out, params, config)                      # keeps mu as the starting point, 20 more E-iterations
    mu_tr = np.stack([t["mu"] for t in out])
    assert np.isfinite(mu_tr).all() and relerr(mu_tr, mu_fit) < 0.2


def test_overlapped_m_and_h_step_is_bit_identical(vl):
    """vem runs the M-step on a second stream under the H-step (independent given the E-step, vlgp/core.py:318-325):
    the result must be bit-for-bit what the sequential order gives, and begin/end must behave when misused."""
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky
    from vlgp_b200._lib import VlgpNativeError

    out = []
    for overlap in (True, False):
        segs, params = _problem(77, 6, 200, 23, 3)
        cfg = _cfg(max_iter=3, min_iter=3, Eniter=5, Mniter=6)
        cfg["overlap_mh"] = overlap
        make_cholesky(segs, params, cfg)
        core.vem(segs, params, cfg)
        out.append((params, np.stack([s["mu"] for s in segs]), np.stack([s["v"] for s in segs])))
    (p1, mu1, v1), (p0, mu0, v0) = out
    for k in ("a", "b", "noise", "da", "db", "omega", "sigma"):
        assert np.array_equal(p1[k], p0[k]), k
    assert np.array_equal(mu1, mu0) and np.array_equal(v1, v0)

    segs, params = _problem(78, 2, 100, 8, 2)
    with core.Session(segs, params) as s:
        assert s.ts.mstep_end() == 0                       # nothing pending: a no-op
        s.ts.mstep_begin(3)
        with pytest.raises(VlgpNativeError):
            s.ts.mstep_begin(3)                            # one at a time
        a_dev = s.eng.pull_params({})["a"]                 # waits for the pending M-step by itself
        assert s.ts.mstep_end() == 0
        assert np.array_equal(a_dev, s.eng.pull_params({})["a"])
        ref_segs, ref_params = _problem(78, 2, 100, 8, 2)
        core.mstep(ref_segs, ref_params, _cfg(Mniter=3))
        assert np.array_equal(a_dev, ref_params["a"])


def test_vem_download_through_prefetched_pinned_blocks(vl, monkeypatch):
    """vem() starts the download of its result behind the last E-step; w and dmu arrive in page-locked blocks that become
    the trial arrays (vlgp_trials_prefetch_state_into / _take).  Same values as the plain download, the reference's
    aliasing (mu, v in place; w, dmu rebound: vlgp/core.py:117-120), blocks recycled once their arrays are gone, and a
    prefetch made stale by a later write of the state is not used."""
    import gc
    from vlgp_b200 import core, engine
    from vlgp_b200.gp import make_cholesky

    pool = engine.get_engine().pinned
    returned = []
    monkeypatch.setattr(pool, "give_back", lambda addr, nbytes, _orig=pool.give_back: (returned.append(addr),
                                                                                       _orig(addr, nbytes))[1])
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VLGP_PREFETCH", mode)
        segs, params = _problem(91, 4, 150, 17, 3)
        cfg = _cfg(max_iter=2, min_iter=2, Eniter=4, Mniter=4)
        make_cholesky(segs, params, cfg)
        mu_before = [s["mu"] for s in segs]
        w_before = [s["w"] for s in segs]
        core.vem(segs, params, cfg)
        assert all(a is b for a, b in zip(mu_before, (s["mu"] for s in segs)))          # in place
        assert not any(a is b for a, b in zip(w_before, (s["w"] for s in segs)))        # rebound
        out[mode] = {k: np.stack([s[k] for s in segs]) for k in ("mu", "v", "w", "dmu")}
        if mode == "1":
            held = segs                                    # keeps the pinned-backed arrays alive
            assert segs[0]["w"].flags.writeable and not segs[0]["w"].flags.owndata
    for k in ("mu", "v", "w", "dmu"):
        assert np.array_equal(out["1"][k], out["0"][k]), k
    # the two blocks go back to the pool with their last view
    assert returned == []
    del held, segs
    gc.collect()
    assert len(returned) == 2 and returned[0] != returned[1]
    # a stale prefetch is ignored: write the state after the prefetch, then pull
    monkeypatch.setenv("VLGP_PREFETCH", "1")
    segs, params = _problem(92, 2, 100, 9, 2)
    cfg = _cfg(Eniter=3)
    make_cholesky(segs, params, cfg)
    with core.Session(segs, params) as s:
        s.ts.estep(3)
        s.ts.prefetch_state(direct=("w", "dmu"))
        s.ts.estep(1)                                      # the state moves on
        assert s.ts.take_prefetched("w") is None
        s.pull(segs)
        ref = s.ts.get_state(("w",))["w"]
    assert np.array_equal(np.concatenate([sg["w"] for sg in segs]), ref)


def test_errors_are_loud(vl):
    from vlgp_b200 import core
    from vlgp_b200._lib import VlgpNativeError

    segs, params = _problem(2, 1, 50, 6, 2)
    del params["cholesky"]
    with pytest.raises(KeyError):
        core.estep(segs, params, _cfg())
    segs[0]["x"] = np.zeros((50, 2, 6))                      # a regressor block that does not match xdim = 1
    with pytest.raises(ValueError):
        core.mstep(segs, params, _cfg())
    eng = __import__("vlgp_b200.engine", fromlist=["get_engine"]).get_engine()
    with pytest.raises(VlgpNativeError):
        eng._ck(eng.lib.vlgp_trials_free(eng.ctx, 12345), "trials_free")
