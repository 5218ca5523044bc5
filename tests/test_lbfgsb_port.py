"""The restated L-BFGS-B (vlgp_b200/csrc/lbfgsb.cuh) against scipy's own routine, evaluation by evaluation.

The reference's H-step calls scipy.optimize.minimize(method="L-BFGS-B") per latent (vlgp/gp.py:100-123); scipy is an
unpinned dependency outside /root/reference, so parity of the port is anchored on running both on identical objectives:
smooth bounded test functions in 1-4 variables, and the H-step's own objective (oracle/vlgp_oracle.py::hstep_objective),
whose gradient is not the derivative of its value, so that its line searches collapse and L-BFGS-B takes its failure
paths (restart from steepest descent, ABNORMAL termination).  CPU only: the optimiser is context-free host code.
"""
import ctypes as C

import numpy as np
import pytest

from vlgp_b200 import _lib

dp = _lib.c_double_p


def _p(a):
    return a.ctypes.data_as(dp)


def run_port(fun, x0, bounds, collapse=0.0, maxls=20):
    lib = _lib.load()
    x0 = np.asarray(x0, dtype=float)
    lo, up = np.ascontiguousarray(bounds[:, 0]), np.ascontiguousarray(bounds[:, 1])
    h = C.c_void_p()
    assert lib.vlgp_lbfgsb_new(len(x0), _p(x0), _p(lo), _p(up), 1e7, 1e-5, maxls, collapse, C.byref(h)) == 0
    x, g, f, task, trace = np.zeros(len(x0)), np.zeros(len(x0)), 0.0, C.c_int(), []
    try:
        while True:
            need = lib.vlgp_lbfgsb_advance(h, f, _p(g), _p(x), C.byref(task))
            assert need >= 0
            if not need:
                break
            f, g = fun(x.copy())
            g = np.ascontiguousarray(g, dtype=float)
            trace.append((x.copy(), f))
        ncol = C.c_int()
        lib.vlgp_lbfgsb_info(h, None, None, None, C.byref(ncol))
    finally:
        lib.vlgp_lbfgsb_free(h)
    return x.copy(), trace, task.value, ncol.value


def run_scipy(fun, x0, bounds, maxls=20):
    """scipy's setulb in reverse communication, exactly as scipy.optimize._lbfgsb_py drives it."""
    from scipy.optimize import _lbfgsb_py as _L

    setulb = _L._lbfgsb.setulb
    it = np.int64 if getattr(_L, "HAS_ILP64", False) else np.int32
    n, m = len(x0), 10
    low, up = np.ascontiguousarray(bounds[:, 0]), np.ascontiguousarray(bounds[:, 1])
    nbd = np.full(n, 2, dtype=it)
    x = np.clip(np.asarray(x0, dtype=float), low, up)
    f, g = np.array(0.0), np.zeros(n)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
    iwa, task, ln_task = np.zeros(3 * n, dtype=it), np.zeros(2, dtype=it), np.zeros(2, dtype=it)
    lsave, isave, dsave = np.zeros(4, dtype=it), np.zeros(44, dtype=it), np.zeros(29)
    trace = []
    while True:
        g = g.astype(np.float64)
        setulb(m, x, low, up, nbd, f, g, 1e7, 1e-5, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)
        if task[0] == 3:
            f, g = fun(x.copy())
            trace.append((x.copy(), f))
        elif task[0] != 1:
            break
    return x.copy(), trace, int(task[0]), int(task[1])


def _same_iterates(ts, tp, rtol, atol):
    return len(ts) == len(tp) and all(np.allclose(a[0], b[0], rtol=rtol, atol=atol) for a, b in zip(ts, tp))


try:
    from scipy.optimize import _lbfgsb_py as _probe

    _probe._lbfgsb.setulb
    HAVE_SETULB = True
except Exception:  # pragma: no cover
    HAVE_SETULB = False

needs_setulb = pytest.mark.skipif(not HAVE_SETULB, reason="scipy's private setulb entry point is not available")


@needs_setulb
def test_smooth_bounded_functions_follow_scipy_evaluation_by_evaluation():
    rng = np.random.default_rng(0)
    nfev = []
    for case in range(120):
        n = int(rng.integers(1, 5))
        A = rng.standard_normal((n, n))
        Q, c, k = A @ A.T + 0.1 * np.eye(n), 2 * rng.standard_normal(n), rng.uniform(0, 2)

        def fun(x):
            return (0.5 * x @ Q @ x - c @ x + k * np.sum(np.cos(1.3 * x)) + 0.1 * np.sum(x ** 4),
                    Q @ x - c - 1.3 * k * np.sin(1.3 * x) + 0.4 * x ** 3)

        b = np.stack([-rng.uniform(0.1, 2, n), rng.uniform(0.1, 2, n)], 1)
        x0 = rng.uniform(-2.5, 2.5, n)
        xs, ts, tk, _ = run_scipy(fun, x0, b)
        xp, tp, tkp, _ = run_port(fun, x0, b)
        assert _same_iterates(ts, tp, 0, 1e-10), case
        assert np.abs(xs - xp).max() < 1e-10
        nfev.append(len(ts))
    assert max(nfev) >= 10


@needs_setulb
def test_ill_conditioned_chain_same_evaluation_counts_and_termination():
    """Chained Rosenbrock, badly scaled, bounds active: dozens of iterations (the limited memory wraps around, m = 10),
    subspace minimisation with projection.  Rounding differences are amplified along such runs, so the iterates are
    compared to 1e-6 and the evaluation counts / termination reasons exactly."""
    rng = np.random.default_rng(1)
    term = {2: 401, 3: 402}            # lbfgsb::Task -> scipy's task[1] of a CONVERGENCE exit
    longest = 0
    for case in range(60):
        n = int(rng.integers(2, 5))
        sc = 10 ** rng.uniform(-1, 1, n)

        def fun(x):
            y = x * sc
            f = np.sum(100 * (y[1:] - y[:-1] ** 2) ** 2 + (1 - y[:-1]) ** 2)
            g = np.zeros(n)
            g[:-1] += -400 * y[:-1] * (y[1:] - y[:-1] ** 2) - 2 * (1 - y[:-1])
            g[1:] += 200 * (y[1:] - y[:-1] ** 2)
            return f, g * sc

        b = np.stack([-rng.uniform(0.5, 3, n) / sc, rng.uniform(0.3, 3, n) / sc], 1)
        x0 = rng.uniform(-2, 2, n) / sc
        xs, ts, tk, tk1 = run_scipy(fun, x0, b)
        xp, tp, tkp, _ = run_port(fun, x0, b)
        assert len(ts) == len(tp), case
        assert tk == 4 and term[tkp] == tk1, case
        assert np.abs(xs - xp).max() < 1e-6 * max(1.0, np.abs(xs).max()), case
        longest = max(longest, len(ts))
    assert longest > 40


def _hstep_problem(seed, omega0):
    """(fun, x0, bounds) per latent for a small H-step after one EM iteration of the oracle."""
    from oracle import vlgp_oracle as orc
    from vlgp_b200.synth import make_trials

    rng = np.random.default_rng(seed)
    N, L, W = 12, 2, 25
    trials = make_trials(4, 100, N, L, seed=seed)
    params = dict(a=0.3 * rng.standard_normal((L, N)), b=np.full((1, N), np.log(0.08)), noise=np.ones(N),
                  omega=np.full(L, omega0), sigma=np.ones(L), likelihood=np.array(["poisson"] * N), zdim=L, ydim=N,
                  xdim=1, rank=25, gp_noise=1e-4, dt=1)
    params["da"], params["db"] = np.zeros_like(params["a"]), np.zeros_like(params["b"])
    segs = []
    for tr in trials:
        for s in range(0, 100, W):
            segs.append(dict(y=tr["y"][s:s + W], x=np.ones((W, 1, N)), mu=0.3 * rng.standard_normal((W, L)),
                             v=np.zeros((W, L)), w=np.zeros((W, L)), dmu=np.zeros((W, L))))
    cfg = orc.default_config(max_iter=1, min_iter=1, Hstep=False, window=W)
    params["cholesky"] = orc.make_cholesky([W], params["omega"], params["sigma"], 25)
    orc.update_w(segs, params, cfg)
    orc.update_v(segs, params, cfg)
    orc.vem(segs, params, cfg)
    mu, w = np.stack([t["mu"] for t in segs]), np.stack([t["w"] for t in segs])
    t = np.arange(W) * 1.0
    bounds = np.log(((1e-3, 1), cfg["omega_bound"], (1e-4 / 2, 2e-4)))
    out = []
    for l in range(L):
        def fun(x, l=l):
            return orc.hstep_objective(x, t, mu[:, :, l].T, w[:, :, l].T)

        out.append((fun, np.log((1.0, omega0, 1e-4)), bounds))
    return out


@needs_setulb
@pytest.mark.parametrize("seed,omega0", [(0, 0.05), (1, 1e-3), (2, 0.01)])
def test_hstep_objective_identical_iterates_and_collapse_rule(seed, omega0):
    """On the reference's own H-step objective the port must request the very same points as scipy (bit for bit: the
    objective is evaluated by the same NumPy code for both), through collapsing line searches, restarts and ABNORMAL
    terminations; with the collapse rule the end point stays within 1e-9 of scipy's and the evaluations drop."""
    saved = 0
    for fun, x0, bounds in _hstep_problem(seed, omega0):
        xs, ts, _, _ = run_scipy(fun, x0, bounds)
        xp, tp, _, _ = run_port(fun, x0, bounds)
        assert len(ts) == len(tp)
        assert all(np.array_equal(a[0], b[0]) for a, b in zip(ts, tp))
        assert np.array_equal(xs, xp)
        xc, tc, _, _ = run_port(fun, x0, bounds, collapse=1e-9)
        assert np.abs(xc - xs).max() <= 1e-9
        saved += len(ts) - len(tc)
    assert saved >= 0


def test_argument_errors_and_limits():
    lib = _lib.load()
    h = C.c_void_p()
    x = np.zeros(5)
    assert lib.vlgp_lbfgsb_new(5, _p(x), _p(x), _p(x), 1e7, 1e-5, 20, 0.0, C.byref(h)) < 0          # n > 4
    assert lib.vlgp_lbfgsb_advance(None, 0.0, None, None, None) < 0
    # a quadratic from inside the box converges on the projected-gradient test
    b = np.array([[-1.0, 1.0]])
    xs, tr, task, _ = run_port(lambda x: (float(x @ x), 2 * x), [0.7], b)
    assert task == 2 and abs(xs[0]) < 1e-5 and len(tr) <= 4
