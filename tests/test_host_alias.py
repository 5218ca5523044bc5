"""The reference updates params["a"] / ["b"] in place through vem, and that identity is observable: params["a"] IS
FactorAnalysis.components_ (vlgp/preprocess.py:20,27) and params["transform"] is the estimator's bound transform, so
transform() on new trials maps them with the fitted loading.  The host code of this package must keep that identity
(CPU-only checks with stand-ins for the device calls)."""
import ctypes as C

import numpy as np
import pytest

from vlgp_b200 import core, preprocess
from vlgp_b200.engine import Engine
from vlgp_b200.util import assign_inplace


def test_assign_inplace_keeps_identity_when_it_can():
    d = {"a": np.zeros((2, 3))}
    keep = d["a"]
    assert assign_inplace(d, "a", np.arange(6.0).reshape(2, 3)) is keep and keep[1, 2] == 5.0
    assign_inplace(d, "a", np.ones((3, 3)))                       # other shape: rebound
    assert d["a"] is not keep and d["a"].shape == (3, 3)
    ro = np.zeros(4)
    ro.flags.writeable = False
    d = {"x": ro}
    assign_inplace(d, "x", np.ones(4))                            # read-only: rebound, the caller's array untouched
    assert d["x"] is not ro and ro.sum() == 0.0
    d = {}
    v = np.ones(2)
    assign_inplace(d, "new", v)
    assert d["new"] is not v and np.array_equal(d["new"], v)      # never aliases the value passed in


class _FakeSet:
    def __init__(self, mom=None):
        self.affine = []
        self.mom = mom

    def latent_affine(self, shift, M):
        self.affine.append((None if shift is None else np.array(shift), None if M is None else np.array(M)))

    def latent_moments(self):
        return self.mom


class _FakeEngine:
    def __init__(self):
        self.pushed = []

    def push_params(self, params, which=()):
        self.pushed.append({k: np.array(params[k]) for k in which})


class _FakeSession:
    alias = None                    # no overlapping windows

    def __init__(self, mom=None):
        self.eng = _FakeEngine()
        self.ts = _FakeSet(mom)


def test_constraints_update_loading_and_bias_in_place():
    rng = np.random.default_rng(0)
    a0 = rng.standard_normal((2, 5))
    for kind in ("fro", 2):
        params = {"a": a0.copy(), "b": np.zeros((1, 5))}
        keep = params["a"]
        s = _FakeSession()
        core._constrain_loading_dev(s, params, {"constrain_loading": kind, "eps": 1e-8})
        assert params["a"] is keep                                             # same array object, new values
        sc = (np.linalg.norm(a0) if kind == "fro" else np.linalg.norm(a0, ord=2, axis=1, keepdims=True)) + 1e-8
        assert np.allclose(keep, a0 / sc) and np.array_equal(s.eng.pushed[0]["a"], keep)
    params = {"a": a0.copy(), "b": np.zeros((1, 5))}
    keep = params["a"]
    core._constrain_loading_dev(_FakeSession(), params, {"constrain_loading": "svd", "eps": 1e-8})
    assert params["a"] is not keep and np.array_equal(keep, a0)                # the reference rebinds here too (:406)

    mean, std, cnt = np.array([0.5, -1.0]), np.array([2.0, 0.25]), 100.0
    mom = (mean * cnt, (std ** 2 + mean ** 2) * cnt, cnt)
    params = {"a": a0.copy(), "b": np.ones((1, 5))}
    ka, kb = params["a"], params["b"]
    s = _FakeSession(mom)
    core._constrain_latent_dev(s, params, {"constrain_latent": "both"})
    assert params["a"] is ka and params["b"] is kb
    assert np.allclose(kb, 1.0 + mean @ a0) and np.allclose(ka, a0 * std[:, None])
    shift, M = s.ts.affine[0]
    assert np.allclose(shift, mean) and np.allclose(M, np.diag(1.0 / std))


def test_pull_params_writes_into_the_arrays_of_the_dict():
    L, N = 2, 3
    dev = {"a": np.arange(6.0).reshape(L, N), "b": np.array([7.0, 8.0, 9.0]), "noise": np.array([.1, .2, .3]),
           "da": np.full((L, N), 0.5), "db": np.array([1.0, 2.0, 3.0])}

    class Lib:
        @staticmethod
        def vlgp_get_params(ctx, a, b, noise, da, db, sigma, omega):
            for ptr, key in ((a, "a"), (b, "b"), (noise, "noise"), (da, "da"), (db, "db")):
                if ptr:
                    C.memmove(ptr, dev[key].ctypes.data, dev[key].nbytes)
            return 0

    eng = Engine.__new__(Engine)
    eng.lib, eng.ctx, eng.L, eng.N, eng.xdim = Lib, None, L, N, 1
    eng._ck = lambda rc, what: None
    params = {"a": np.zeros((L, N)), "b": np.zeros((1, N)), "noise": np.zeros(N)}
    ka, kb, kn = params["a"], params["b"], params["noise"]
    eng.pull_params(params)
    assert params["a"] is ka and np.array_equal(ka, dev["a"])
    assert params["b"] is kb and np.array_equal(kb[0], dev["b"])
    assert params["noise"] is not kn and np.array_equal(params["noise"], dev["noise"])     # rebound like the reference
    assert params["da"].shape == (L, N) and params["db"].shape == (1, N)


def test_fitted_loading_reaches_the_factor_analysis_map():
    """initialize() binds params['a'] to FactorAnalysis.components_; an in-place update must change what
    params['transform'] returns, a rebinding must not (this is the reference's behaviour, pinned end to end against
    the reference in tests/test_oracle_golden.py::test_transform_new_trials_pipeline)."""
    rng = np.random.default_rng(1)
    trials = [{"y": rng.poisson(0.3, size=(120, 8)).astype(float)} for _ in range(5)]
    config = preprocess.get_config()
    params = preprocess.get_params(trials, 2, omega_bound=config["omega_bound"])
    np.random.seed(0)
    preprocess.initialize(trials, params, config)
    y = trials[0]["y"]
    z0 = params["transform"](y)
    assign_inplace(params, "a", params["a"] * 0.5)
    z1 = params["transform"](y)
    assert not np.allclose(z0, z1)


# ----------------------------------------------------------------------------------------------------------------------
# host packing helpers of core.Session (pure host code, no engine)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_c_helper", [True, False])
def test_regressor_check_scans_owners_not_segments(monkeypatch, use_c_helper):
    """The tuned kernels treat the all-ones bias regressor (xdim == 1) as a per-neuron constant; anything else goes
    through the general-regressor path (csrc/regress.cu).  Segments are views of their trial's x: the classification
    must accept them, scan each owning array once, and still notice every way a regressor can be something else."""
    from vlgp_b200 import core

    if not use_c_helper:
        monkeypatch.setattr(core, "_fastpack", None)
    elif core._fastpack is None:
        pytest.skip("_fastpack not built")
    T, N = 120, 7
    P = dict(xdim=1)
    x = np.ones((T, 1, N))
    segs = [dict(x=x[s:s + 40]) for s in range(0, T, 40)] + [dict(), dict(x=None), dict(x=np.ones((30, 1, N)))]
    scans = []
    real = core._all_ones
    monkeypatch.setattr(core, "_all_ones", lambda a: scans.append(a.shape) or real(a))
    assert core._bias_only(segs, P)
    if use_c_helper:
        assert sorted(scans) == [(30, 1, N), (T, 1, N)]          # one scan per owner, not per segment
    bad = np.ones((T, 1, N))
    bad[77, 0, 3] = 0.5
    assert not core._bias_only(segs + [dict(x=bad[40:80])], P)
    assert core._bias_only(segs + [dict(x=bad[:40])], P)         # the view itself is all ones even if its owner is not
    assert not core._bias_only([dict(x=np.ones((T, 2, N)))], P)
    assert not core._bias_only([dict(x=np.ones((T, N)))], P)
    assert not core._bias_only(segs, dict(xdim=2))               # history > 1: b has xdim rows
    # (an array verified once is remembered as all ones for as long as it lives, hence fresh arrays per case)
    wide = np.ones((T, 2, N))
    assert core._bias_only([dict(x=wide[:, :1, :])], P)          # a (T, 1, N) view of a wider array of ones
    wide = np.ones((T, 2, N))
    wide[5, 1, 0] = 3.0
    assert core._bias_only([dict(x=wide[:, :1, :])], P)          # ... judged by its own entries
    wide = np.ones((T, 2, N))
    wide[5, 0, 0] = 3.0
    assert not core._bias_only([dict(x=wide[:, :1, :])], P)


def test_row_views_equal_np_split():
    from vlgp_b200 import core

    for lengths in ([50] * 6, [50, 120, 64, 1, 200], [7]):
        lengths = np.asarray(lengths, dtype=np.int32)
        starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
        a = np.arange(int(lengths.sum()) * 3, dtype=float).reshape(-1, 3)
        views = core._row_views(a, starts, lengths)
        ref = np.split(a, [int(s) for s in starts[1:]])
        assert len(views) == len(ref)
        for v, r in zip(views, ref):
            assert v.shape == r.shape and np.shares_memory(v, a) and np.array_equal(v, r)
            assert v.flags.c_contiguous and v.flags.writeable
