import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def relerr(x, ref):
    x = np.asarray(x, dtype=float)
    ref = np.asarray(ref, dtype=float)
    return float(np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300))


def use_oracle_prior_factors(monkeypatch):
    """Prior factors from the oracle's ichol_gauss (bit-identical to the reference's, tests/test_oracle_golden.py) in
    place of the device's, uploaded with set_cholesky.  The device's own factorisation has the reference's pivots on the
    golden grid (test_ichol_pivots_and_factor) but breaks EXACT ties between mirror-image pivots by its own rounding on
    ~10 % of random (n, omega): a whole-pipeline comparison at 1e-7 has to take that step out."""
    from oracle import vlgp_oracle as orc
    from vlgp_b200 import api, core, gp
    from vlgp_b200.engine import get_engine

    def session_make_cholesky(self, params):
        self.eng.push_params(params, which=("sigma", "omega"))
        lengths = sorted(set(int(t) for t in self.ts.lengths.tolist()))
        chol = orc.make_cholesky(lengths, params["omega"], params["sigma"], params["rank"])
        params["cholesky"] = {t: chol[t] for t in lengths}
        for t in lengths:
            self.ts.set_cholesky(t, chol[t])
        self.have_factors = True

    def make_cholesky(trials, params, config=None):
        eng = get_engine()
        eng.ensure_model(params)
        eng.push_params(params, which=("sigma", "omega"))
        lengths = sorted({int(tr["y"].shape[0]) for tr in trials})
        params["cholesky"] = orc.make_cholesky(lengths, params["omega"], params["sigma"], params["rank"])

    monkeypatch.setattr(core.Session, "make_cholesky", session_make_cholesky)
    monkeypatch.setattr(gp, "make_cholesky", make_cholesky)
    monkeypatch.setattr(api, "make_cholesky", make_cholesky)


def inject_hyperparameter_trajectory(monkeypatch, omega_traj, sigma_traj):
    """Replace the H-step's optimiser by one that returns the reference's own result of each EM iteration (recorded by
    oracle/make_golden.py::golden_fit through a callback).  Everything around the optimiser -- new prior factors from
    the new omega (vlgp/gp.py:94-97), the following E- and M-steps, the final inference -- runs as usual, so a whole
    default fit() can be compared with the reference without L-BFGS-B's sensitivity to the last digits of its objective
    (DESIGN.md section 5)."""
    import numpy as np
    from vlgp_b200 import gp

    state = {"it": 0}

    def injected(s, params, config):
        k = state["it"]
        state["it"] += 1
        params["sigma"] = np.array(sigma_traj[k], dtype=float)
        params["omega"] = np.array(omega_traj[k], dtype=float)
        config.setdefault("hstep_nfev", []).append([0] * len(omega_traj[k]))
        s.make_cholesky(params)

    monkeypatch.setattr(gp, "_optimize_dev_impl", injected)
    return state
