"""Lane-level (32 lanes, NumPy) emulation of the FP64 tensor-core block sweep of the H-step kernel: the products that the
HALF_LAST instantiation leaves out multiply exact zeros, so its result is the full sweep's bit for bit and equals
-B^-1 (scripts/dmma_half_last_emulation.py; hstep_dmma.cu)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_half_last_sweep_is_identical_to_the_full_sweep(capsys):
    spec = importlib.util.spec_from_file_location("half_last", os.path.join(ROOT, "scripts", "dmma_half_last_emulation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()                                  # asserts inside: zero operands, identical tiles, inverse to 1e-12
    out = capsys.readouterr().out
    assert "W=50 NB=7: DMMA 378 -> 351" in out


@pytest.mark.parametrize("W,q", [(50, 2), (49, 1), (18, 2), (17, 1)])
def test_bordered_schur_identities_of_the_hstep_kernel(W, q):
    """The algebra behind hstep_segment_schur_kernel (csrc/hstep_dmma.cu), in NumPy: with the q = 1 or 2 border rows of
    B = I + d K d eliminated first (X = B12 B22^-1, S = B11 - X B21), tr(B^-1) and tr(B^-1 C), C = d dK d, are
    element-wise sums over S^-1 with rank-modified coefficients -- no product with the border is formed.  Uses the
    kernel's own vectors (u, v, x1, x2, G, H) and its half-symmetrised coefficient, against the direct inverse."""
    rng = np.random.default_rng(W)
    t = np.arange(W, dtype=float)
    omega, eps = 7e-3, 1e-4
    D2 = (t[:, None] - t[None, :]) ** 2
    Ks = np.exp(-omega * D2)
    K = Ks + eps * np.eye(W)
    dK = -Ks * D2 * omega
    d = np.sqrt(rng.uniform(0.0, 3.0, W))
    B = np.eye(W) + d[:, None] * K * d[None, :]
    C = d[:, None] * dK * d[None, :]
    Binv = np.linalg.inv(B)
    tr_ref, pd_ref = np.trace(Binv), np.sum(Binv * C)

    WC = W - q
    a, b = WC, WC + 1
    da, db = d[a], (d[b] if q == 2 else 0.0)
    Kab = K[a, b] if q == 2 else 0.0
    Kbb = K[b, b] if q == 2 else 0.0
    b00, b01, b11 = 1.0 + da * K[a, a] * da, da * Kab * db, 1.0 + db * Kbb * db
    det = b00 * b11 - b01 * b01
    p, qq, rr = b11 / det, -b01 / det, b00 / det
    c01 = da * (dK[a, b] if q == 2 else 0.0) * db
    dc = d[:WC]
    u = dc * K[:WC, a] * da
    v = dc * (K[:WC, b] if q == 2 else 0.0) * db
    x1, x2 = p * u + qq * v, qq * u + rr * v
    g = da * dK[a, :WC] * dc
    h = db * (dK[b, :WC] if q == 2 else 0.0) * dc
    G, H = g - 0.5 * c01 * x2, h - 0.5 * c01 * x1
    S = np.eye(WC) + dc[:, None] * K[:WC, :WC] * dc[None, :] - np.outer(u, x1) - np.outer(v, x2)
    assert np.allclose(S, S.T, atol=1e-13)
    Sinv = np.linalg.inv(S)
    kap = (dc[:, None] * dK[:WC, :WC] * dc[None, :] - np.outer(x1, G) - np.outer(x2, H)
           - np.outer(G, x1) - np.outer(H, x2))
    tau = np.outer(x1, x1) + np.outer(x2, x2)
    pd = np.sum(Sinv * kap) + 2.0 * qq * c01
    tr = np.trace(Sinv) + np.sum(Sinv * tau) + p + (rr if q == 2 else 0.0)
    assert abs(tr - tr_ref) < 1e-11 * abs(tr_ref)
    assert abs(pd - pd_ref) < 1e-10 * max(abs(pd_ref), 1.0)
