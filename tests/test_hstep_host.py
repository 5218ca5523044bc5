"""Host side of the H-step (vlgp_b200/gp.py) on the CPU with a stand-in for the device objective: the lockstep driver
of scipy's L-BFGS-B routine must produce, for every latent, exactly the iterate sequence -- hence the same optimum, final
value and evaluation count -- as the reference's `scipy.optimize.minimize(method="L-BFGS-B", jac=True, bounds=...)`
call (vlgp/gp.py:100-123), and `_optimize_dev` must apply the reference's acceptance rule (vlgp/gp.py:84-97)."""
import threading

import numpy as np
import pytest
from scipy.optimize import minimize

from vlgp_b200 import gp

GP_NOISE = 1e-4
BOUNDS = ((1e-3, 1), (5e-4, 5e-2), (GP_NOISE / 2, GP_NOISE * 2))
MASK = np.array([0, 1, 0])


class FakeSet:
    """ll(omega) per latent: smooth, with the optimum inside, on, or beyond the omega bounds; evaluation counter;
    `bad_once` makes the first evaluation of a latent report "K not positive definite" (info = 1)."""

    def __init__(self, centres, bad_once=()):
        self.c = np.log(np.asarray(centres, dtype=float))
        self.calls = 0
        self.rounds = 0
        self.bad = set(bad_once)
        self.lock = threading.Lock()

    def _one(self, l, hyper):
        x = np.log(hyper[1])
        d = x - self.c[l]
        ll = -(12.0 * d ** 2 + 3.0 * d ** 4 + 0.2 * np.sin(5.0 * x)) * (1.0 + 0.1 * l) - 3.0 * hyper[0]
        dll = -(24.0 * d + 12.0 * d ** 3 + np.cos(5.0 * x)) * (1.0 + 0.1 * l)
        return ll, dll

    def hstep_prepare(self):
        pass

    def hstep_objective(self, l, hyper):
        with self.lock:
            self.calls += 1
            if l in self.bad:
                self.bad.discard(l)
                return 0.0, 0.0, 1
        ll, dll = self._one(l, np.asarray(hyper, dtype=float))
        return float(ll), float(dll), 0

    def hstep_objective_batch(self, latents, hypers):
        self.rounds += 1
        out = [self.hstep_objective(int(l), h) for l, h in zip(latents, np.asarray(hypers, dtype=float))]
        return (np.array([o[0] for o in out]), np.array([o[1] for o in out]), np.array([o[2] for o in out], dtype=np.int32))


def _reference_run(ts, l, initial):
    """What the reference does for one latent: minimize over log-hyperparameters with the masked gradient."""
    fun = gp._objective(lambda lat, h: ts.hstep_objective(lat, h), l, MASK)
    res = minimize(fun, np.log(initial), jac=True, bounds=np.log(BOUNDS), method="L-BFGS-B")
    return res.x, float(res.fun), int(res.nfev)


CENTRES = [3e-3, 2e-2, 1e-4, 0.5, 7e-3]          # inside, inside, below the lower bound, above the upper bound, inside
INITIALS = [(1.0, 0.05, GP_NOISE), (0.8, 0.01, GP_NOISE), (1.0, 0.02, GP_NOISE), (1.0, 1e-3, GP_NOISE), (0.5, 0.03, GP_NOISE)]


def test_lockstep_driver_equals_scipy_minimize_bit_for_bit():
    ts = FakeSet(CENTRES)
    lock = gp._lockstep_lbfgsb(ts, list(range(5)), INITIALS, BOUNDS, MASK)
    assert lock is not None, "scipy's private setulb entry point is missing: the threaded fallback would be used"
    for l in range(5):
        x_ref, f_ref, n_ref = _reference_run(FakeSet(CENTRES), l, INITIALS[l])
        x, f, n = lock[l]
        assert np.array_equal(x, x_ref), (l, x, x_ref)
        assert f == f_ref and n == n_ref, (l, f, f_ref, n, n_ref)
    # one batched device call per round: as many rounds as the slowest latent needs evaluations
    assert ts.rounds == max(n for _, _, n in lock)
    assert ts.calls == sum(n for _, _, n in lock)


def test_single_latent_driver_and_not_pd_retry():
    ts = FakeSet(CENTRES, bad_once={1})
    x, f, n = gp.optimze1d(ts, 1, INITIALS[1], BOUNDS, MASK)
    ref = FakeSet(CENTRES, bad_once={1})
    x_ref, f_ref, n_ref = _reference_run(ref, 1, INITIALS[1])
    assert np.array_equal(np.log(x), x_ref) and f == f_ref and n == n_ref
    assert ts.calls == ref.calls == n + 1            # the retry costs one extra device evaluation, not an nfev
    # the lockstep driver retries inside a round as well
    ts2 = FakeSet(CENTRES, bad_once={0, 4})
    lock = gp._lockstep_lbfgsb(ts2, [0, 4], [INITIALS[0], INITIALS[4]], BOUNDS, MASK)
    assert lock is not None and np.isfinite([r[1] for r in lock]).all()


class FakeSession:
    def __init__(self, ts):
        self.ts = ts
        self.cholesky_calls = 0

    def make_cholesky(self, params):
        self.cholesky_calls += 1
        params["cholesky"] = {"made_with": (params["omega"].copy(), params["sigma"].copy())}


@pytest.mark.parametrize("mode", ["lockstep", "threads", "sequential"])
def test_optimize_applies_the_reference_acceptance_rule(mode, monkeypatch):
    for var in ("VLGP_SEQUENTIAL_HSTEP", "VLGP_THREADED_HSTEP"):
        monkeypatch.delenv(var, raising=False)
    if mode == "threads":
        monkeypatch.setenv("VLGP_THREADED_HSTEP", "1")
    elif mode == "sequential":
        monkeypatch.setenv("VLGP_SEQUENTIAL_HSTEP", "1")
    s = FakeSession(FakeSet(CENTRES))
    omega0 = np.array([x[1] for x in INITIALS])
    params = {"zdim": 5, "sigma": np.sqrt([x[0] for x in INITIALS]), "omega": omega0.copy(), "gp_noise": GP_NOISE}
    config = {"omega_bound": BOUNDS[1]}
    gp._optimize_dev(s, params, config)
    want = []
    for l in range(5):
        x_ref, _, n_ref = _reference_run(FakeSet(CENTRES), l, INITIALS[l])
        want.append((np.exp(x_ref), n_ref))
    for l in (0, 1, 4):          # optimum strictly inside the bounds: accepted
        assert params["omega"][l] == want[l][0][1]
    for l in (2, 3):             # optimiser ran into a bound: omega keeps its previous value (vlgp/gp.py:91-92)
        assert np.any(np.isclose(want[l][0][1], BOUNDS[1]))
        assert params["omega"][l] == omega0[l]
    assert np.array_equal(params["sigma"], np.sqrt([w[0][0] for w in want]))        # sigma^2 never moves (mask)
    assert config["hstep_nfev"] == [[w[1] for w in want]]
    assert s.cholesky_calls == 1 and "made_with" in params["cholesky"]              # vlgp/gp.py:97
