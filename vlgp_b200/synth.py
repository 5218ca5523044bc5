"""Deterministic synthetic spike trains of the shapes named in BASELINE.json (host-side, NumPy only).

Mirrors the recipe in SURVEY.md section 8(d) / BASELINE.md section 3 (itself modelled on the reference's
tests/test_api.py:4-28 and notebook/tutorial.ipynb:209-216): sinusoidal latents with random phases, a random
loading with |a| in [0.5, 1), bias log(0.05), Poisson counts.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_trials", "lorenz_latents", "CONFIGS"]

# name -> (n_trials, T or (Tmin, Tmax), neurons, latents, dtype)
CONFIGS = {
    "tutorial": dict(n_trials=10, T=200, N=30, L=3, dtype="f64"),
    "config2": dict(n_trials=256, T=1000, N=100, L=5, dtype="f64"),
    "config3": dict(n_trials=256, T=(500, 2000), N=200, L=10, dtype="f32"),
    "config4": dict(n_trials=1024, T=1000, N=100, L=5, dtype="f64"),
    "config5": dict(n_trials=512, T=2000, N=150, L=3, dtype="f64", latents="lorenz"),
}


def lorenz_latents(n_trials, T, rng, dt=5e-3, skip=2000):
    """Euler-integrated Lorenz attractor, z-scored, reshaped to (n_trials, T, 3).  Same construction as the reference's
    notebook/tutorial.ipynb:143-147 around vlgp/simulation.py:108-151 (s=10, r=28, b=2.667)."""
    n = skip + n_trials * T
    s, r, b = 10.0, 28.0, 2.667
    out = np.empty((n, 3))
    p = rng.random(3)
    for i in range(n):
        dx = s * (p[1] - p[0])
        dy = p[0] * (r - p[2]) - p[1]
        dz = p[0] * p[1] - b * p[2]
        p = p + dt * np.array([dx, dy, dz])
        out[i] = p
    out = out[skip:]
    out = (out - out.mean(axis=0)) / out.std(axis=0)
    return out.reshape(n_trials, T, 3)


def make_trials(n_trials, T, N, L, seed=0, latents="sine", first_id=0, dtype=float):
    """Return a list of ``{"y": (T_i, N) counts, "ID": i}`` trial dicts.

    ``T`` may be an int or a (low, high) pair for unequal lengths (``rng.integers(low, high + 1)`` per trial).
    The generator state depends only on ``seed`` so every rank can build the same global data set and slice it.
    """
    rng = np.random.default_rng(seed)
    a = 0.5 * (rng.random((L, N)) + 1.0) * np.sign(rng.standard_normal((L, N)))
    b = np.log(0.05)
    if isinstance(T, (tuple, list)):
        lengths = rng.integers(T[0], T[1] + 1, size=n_trials)
    else:
        lengths = np.full(n_trials, int(T))
    lor = None
    if latents == "lorenz":
        assert L == 3 and len(set(lengths.tolist())) == 1
        lor = lorenz_latents(n_trials, int(lengths[0]), rng)
    trials = []
    freq = 2.0 * np.pi * (np.arange(L) + 1.0) / 250.0
    for i in range(n_trials):
        Ti = int(lengths[i])
        if lor is not None:
            z = lor[i]
        else:
            phase = rng.uniform(0.0, 2.0 * np.pi, L)
            z = np.sin(np.arange(Ti)[:, None] * freq[None, :] + phase[None, :])
        y = rng.poisson(np.exp(z @ a + b)).astype(dtype)
        trials.append({"y": y, "ID": first_id + i})
    return trials
