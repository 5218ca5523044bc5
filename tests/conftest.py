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
