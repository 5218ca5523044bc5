"""The package's HOST code (api.fit / transform, core.vem + Session, gp's L-BFGS-B driver) driven end to end on the CPU
with the oracle-backed stand-in engine of tests/oracle_engine.py, against the golden vectors of the unmodified
reference.  What this pins without a GPU: the order of the calls fit makes, which arrays of the trial / params dicts
are updated in place and which are rebound, the M-step / H-step overlap's independence assumption, config['runtime'],
callbacks, and transform() on new trials."""
import copy

import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle_engine import install


def test_fit_and_transform_match_the_reference(monkeypatch):
    import vlgp_b200 as vlgp
    from vlgp_b200.synth import make_trials

    eng = install(monkeypatch)
    g, gx = load_golden("fit_fixed_omega"), load_golden("api_extras")
    trials = make_trials(10, 200, 30, 3, seed=0)
    ys = [t["y"] for t in trials]
    np.random.seed(0)
    res = vlgp.fit(trials, 3, max_iter=3, min_iter=3, Hstep=False)
    assert res["trials"] is trials and all(t["y"] is y for t, y in zip(trials, ys))
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in trials]), g[k]) < 1e-9, k
    for k in ("a", "b", "noise"):
        assert relerr(res["params"][k], g[k]) < 1e-9, k
    rt = res["config"]["runtime"]
    assert rt["it"] == 3 and all(len(rt[k]) == 3 for k in ("e_elapsed", "m_elapsed", "h_elapsed", "em_elapsed"))
    assert set(res["params"]["cholesky"]) == {200} and "initial" in res["params"]
    # the call sequence of vlgp/api.py:49-71
    names = [n for n, _ in eng.log if n in ("make_cholesky", "update_w", "update_v", "estep", "mstep", "mstep_begin")]
    assert names[:3] == ["make_cholesky", "update_w", "update_v"]
    assert names[-4:] == ["make_cholesky", "update_w", "update_v", "estep"] and eng.log[-1][0] != "mstep"
    assert [d for n, d in eng.log if n == "estep"] == [25, 25, 25, 3]          # infer: Eniter := max_iter
    # transform() of unseen trials starts from the FactorAnalysis map evaluated with the FITTED loading
    new = make_trials(3, 200, 30, 3, seed=77)
    vlgp.transform(new, res["params"], res["config"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in new]), gx["new_" + k]) < 1e-9, k


def test_fit_with_hstep_matches_the_reference(monkeypatch):
    import vlgp_b200 as vlgp
    from vlgp_b200.synth import make_trials

    eng = install(monkeypatch)
    g = load_golden("fit_tutorial")
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = vlgp.fit(trials, 3, max_iter=3, min_iter=3)
    # the oracle's ichol is bit-identical to the reference's, so unlike on the device there are no pivot-tie flips:
    # what is left is L-BFGS-B's sensitivity to the last digits of its objective
    assert relerr(res["params"]["omega"], g["omega"]) < 1e-6
    for k in ("a", "b"):
        assert relerr(res["params"][k], g[k]) < 1e-6, k
    assert relerr(np.stack([t["mu"] for t in trials]), g["mu"]) < 1e-6
    # M-step begun before the H-step and ended after it, every iteration; the H-step ends with new prior factors
    seq = [n for n, _ in eng.log if n in ("mstep_begin", "mstep_end", "hstep_prepare", "mstep")]
    assert seq == ["mstep_begin", "hstep_prepare", "mstep_end"] * 3
    assert len(res["config"]["hstep_nfev"]) == 3 and all(len(x) == 3 for x in res["config"]["hstep_nfev"])


def test_default_fit_with_the_reference_omega_trajectory(monkeypatch):
    """The DEFAULT fit (H-step on) with the reference's omega of every iteration injected in place of the optimiser's:
    the host code's handling of everything around it (factors rebuilt from the new omega, M-step overlapped with the
    H-step, final inference with the final factors) equals the reference to rounding."""
    import vlgp_b200 as vlgp
    from conftest import inject_hyperparameter_trajectory
    from vlgp_b200.synth import make_trials

    install(monkeypatch)
    g = load_golden("fit_tutorial")
    state = inject_hyperparameter_trajectory(monkeypatch, g["omega_traj"], g["sigma_traj"])
    trials = make_trials(10, 200, 30, 3, seed=0)
    np.random.seed(0)
    res = vlgp.fit(trials, 3, max_iter=3, min_iter=3)
    assert state["it"] == 3
    assert np.array_equal(res["params"]["omega"], g["omega"])
    for k in ("a", "b", "noise"):
        assert relerr(res["params"][k], g[k]) < 1e-9, k
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in res["trials"]]), g[k]) < 1e-9, k


def test_vem_aliasing_callbacks_and_sequential_order(monkeypatch):
    from vlgp_b200 import core, preprocess
    from vlgp_b200.gp import make_cholesky
    from vlgp_b200.synth import make_trials
    from vlgp_b200.util import cut_trials

    install(monkeypatch)
    out = {}
    for overlap in (True, False):
        trials = make_trials(3, 100, 8, 2, seed=4)
        config = preprocess.get_config(max_iter=2, min_iter=2, Eniter=4, Mniter=3)
        config["overlap_mh"] = overlap
        params = preprocess.get_params(trials, 2, omega_bound=config["omega_bound"])
        np.random.seed(1)
        preprocess.initialize(trials, params, config)
        preprocess.fill_params(params)
        preprocess.fill_trials(trials)
        segs = cut_trials(trials, params, config)
        preprocess.fill_trials(segs)
        make_cholesky(segs, params, config)
        a_id, b_id = id(params["a"]), id(params["b"])
        mu_ids = [id(s["mu"]) for s in segs]
        w_ids = [id(s["w"]) for s in segs]
        seen = []
        config["callbacks"] = [lambda tr, pa, co: seen.append((co["runtime"]["it"], tr[0]["mu"].copy(), pa["a"].copy()))]
        core.vem(segs, params, config)
        # segments are views of their trial: mu / v are written through them in place, w / dmu are rebound
        assert [id(s["mu"]) for s in segs] == mu_ids
        assert np.shares_memory(segs[0]["mu"], trials[0]["mu"]) and np.array_equal(trials[0]["mu"][:50], segs[0]["mu"])
        assert [id(s["w"]) for s in segs] != w_ids
        assert id(params["a"]) == a_id and id(params["b"]) == b_id             # updated in place (FactorAnalysis alias)
        assert [it for it, _, _ in seen] == [1, 2]
        assert np.array_equal(seen[-1][1], segs[0]["mu"]) and np.array_equal(seen[-1][2], params["a"])
        out[overlap] = (copy.deepcopy(params), np.concatenate([s["mu"] for s in segs]))
    for k in ("a", "b", "noise", "omega", "sigma", "da", "db"):
        assert np.array_equal(out[True][0][k], out[False][0][k]), k            # overlap = a pure reordering
    assert np.array_equal(out[True][1], out[False][1])


def test_vem_three_iterations_through_the_host_code(monkeypatch):
    from vlgp_b200 import core
    from vlgp_b200.gp import make_cholesky
    from vlgp_b200.preprocess import get_config

    install(monkeypatch)
    g = load_golden("vem")
    n, W, N = g["in_mu"].shape[0], g["y"].shape[1], g["y"].shape[2]
    segs = [dict(y=g["y"][i].astype(float), x=np.ones((W, 1, N)), **{k: g["in_" + k][i].copy() for k in ("mu", "v", "w", "dmu")})
            for i in range(n)]
    a = g["in_a"].copy()
    params = dict(a=a, b=g["in_b"].copy(), noise=g["in_noise"].copy(), omega=g["in_omega"].copy(),
                  sigma=g["in_sigma"].copy(), da=np.zeros_like(a), db=np.zeros_like(g["in_b"]),
                  likelihood=np.array(["poisson"] * N), zdim=a.shape[0], ydim=N, xdim=1, rank=50, gp_noise=1e-4, dt=1)
    cfg = get_config(max_iter=3, min_iter=3)
    make_cholesky(segs, params, cfg)
    core.vem(segs, params, cfg)
    assert cfg["runtime"]["it"] == int(g["n_it"])
    assert relerr(params["omega"], g["out_omega"]) < 1e-6
    for k in ("a", "b"):
        assert relerr(params[k], g["out_" + k]) < 1e-7, k
    assert relerr(np.stack([s["mu"] for s in segs]), g["out_mu"]) < 1e-6


def _fit_option_cases():
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)            # defines the case table; touches the reference only inside main()
    return mod


@pytest.mark.parametrize("case", ["mixed_lik", "user_a_b", "user_omega_sigma", "latent_both_map", "loading_svd",
                                  "tol_early_stop", "window25"])
def test_fit_keyword_arguments_match_the_reference(monkeypatch, case):
    """fit() under non-default keyword arguments -- mixed likelihoods, user-supplied loading / bias / GP
    hyperparameters, latent constraint + MAP, the 'svd' loading constraint (which REBINDS every segment's mu in the
    reference, so the trials keep their initial mu until the final infer), early stop on tol, another window -- through
    the package's host code over the oracle stand-in, against whole-fit outputs of the reference."""
    import vlgp_b200 as vlgp

    mg = _fit_option_cases()
    install(monkeypatch)
    g = load_golden("fit_options")
    N, L, cases = mg.fit_option_cases()
    kw = cases[case]
    trials = mg.fit_option_trials(N, L, kw)
    np.random.seed(0)
    res = vlgp.fit(trials, L, **copy.deepcopy(kw))
    p = case + "/"
    assert res["config"]["runtime"]["it"] == int(g[p + "n_it"])
    for k in ("mu", "v", "w"):
        assert relerr(np.stack([t[k] for t in trials]), g[p + k]) < 1e-8, k
    for k in ("a", "b", "noise", "omega", "sigma"):
        assert relerr(res["params"][k], g[p + k]) < 1e-8, k


@pytest.mark.parametrize("case", ["default", "latent_both_no_hstep", "row_norm_loading"])
def test_fit_on_overlapping_aliased_windows_matches_the_reference(monkeypatch, case):
    """Trial lengths that are not multiples of the window: the reference's segments are overlapping VIEWS, processed
    in order with in-place updates of the shared bins.  core.py reproduces that exactly when the engine offers the
    three row-level operations (_Aliasing); the oracle stand-in does, and the whole fit equals the reference's."""
    import vlgp_b200 as vlgp

    mg = _fit_option_cases()
    eng = install(monkeypatch)
    g = load_golden("fit_overlap")
    trials = mg.fit_overlap_trials()
    np.random.seed(0)
    res = vlgp.fit(trials, 2, **copy.deepcopy(mg.FIT_OVERLAP_CASES[case]))
    p = case + "/"
    for k in ("a", "b", "noise", "omega", "sigma"):
        assert relerr(res["params"][k], g[p + k]) < 1e-9, k
    for k in ("mu", "v", "w"):
        assert relerr(np.concatenate([t[k] for t in trials]), g[p + k]) < 1e-9, k
    assert any(n == "estep" and isinstance(d, tuple) for n, d in eng.log)          # the level-by-level E-step ran


def test_hstep_on_unequal_lengths_raises_like_the_reference(monkeypatch):
    """window=None keeps the trials uncut; with the H-step on the reference stacks their mu / w (vlgp/gp.py:77-80) and
    numpy raises ValueError for unequal lengths.  Same exception type here, before anything is launched; with
    Hstep=False the same call goes through (the E- and M-step take any lengths)."""
    import vlgp_b200 as vlgp
    from vlgp_b200.synth import make_trials

    install(monkeypatch)
    trials = make_trials(1, 80, 6, 2, seed=1) + make_trials(1, 120, 6, 2, seed=2)
    np.random.seed(0)
    with pytest.raises(ValueError, match="same shape"):
        vlgp.fit(trials, 2, max_iter=1, min_iter=1, window=None)
    install(monkeypatch)
    trials = make_trials(1, 80, 6, 2, seed=1) + make_trials(1, 120, 6, 2, seed=2)
    np.random.seed(0)
    res = vlgp.fit(trials, 2, max_iter=1, min_iter=1, window=None, Hstep=False)
    assert res["config"]["runtime"]["it"] == 1 and [t["mu"].shape for t in trials] == [(80, 2), (120, 2)]
