"""CPU oracle for the vLGP variational-EM hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package ``vlgp_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of ``bench.py`` may import this module.

It is a NumPy/SciPy *restatement* (not a copy) of the reference algorithm in catniplab/vlgp, written against flat
arrays so every step can be compared with the CUDA path on identical inputs.  Each function cites the reference
lines it restates (paths relative to /root/reference).  Parity pin: ``tests/test_oracle_golden.py`` checks every function
here against golden vectors produced by the unmodified reference (``oracle/make_golden.py`` via ``oracle/ref_shim.py``),
and -- when /root/reference is present -- against the live reference.

Conventions: N neurons, L latents, r rank of the prior factor, T bins of one trial (or segment).  ``poisson`` is a
boolean mask over neurons (True = Poisson link, False = Gaussian), restating ``params["likelihood"]``.
All arithmetic is float64, like the reference.
"""
from __future__ import annotations

import math
import time

import numpy as np
from scipy.linalg import cho_factor, cho_solve, LinAlgError

__all__ = [
    "trunc_exp", "ichol_gauss", "make_cholesky", "estep_trial", "estep", "update_w", "update_v", "mstep",
    "se_kernel", "posterior_cov", "elbo", "hstep_objective", "hstep", "constrain_loading", "constrain_latent",
    "vem", "cut_trial_starts", "default_config",
]


# --------------------------------------------------------------------------------------------------------------------
# elementwise helpers
# --------------------------------------------------------------------------------------------------------------------
def trunc_exp(x, bound=10.0):
    """exp(min(x, bound)) -- vlgp/math.py:24-38."""
    return np.exp(np.minimum(x, bound))


def _clip_inplace(a, bound):
    """Symmetric in-place clip -- vlgp/util.py:446-454."""
    np.clip(a, -bound, bound, out=a)


def _spd_solve(A, B):
    """Solve A X = B for SPD A via LAPACK potrf/potrs (same driver pair as ``posv`` used by
    ``scipy.linalg.solve(..., sym_pos=True)`` at vlgp/core.py:89,110,193,211,226,230,465)."""
    return cho_solve(cho_factor(A, lower=True, check_finite=False), B, check_finite=False)


# --------------------------------------------------------------------------------------------------------------------
# prior factor
# --------------------------------------------------------------------------------------------------------------------
def ichol_gauss(n, omega, r, dt=1.0, tol=1e-6, return_pivots=False):
    """Pivoted incomplete Cholesky of K_ij = exp(-omega (x_i-x_j)^2), x = dt*arange(n), without forming K.

    Restates vlgp/math.py:76-126: greedy pivot = first arg-max of the residual diagonal (in the current permuted
    order); stop after r columns or once the residual trace drops to tol*n; the residual diagonal is recomputed from
    scratch as 1 - sum_k G_jk^2 after every column; rows are returned in natural order; unused columns stay zero.
    """
    n = int(n)
    pos = np.arange(n) * float(dt)
    perm = np.arange(n)
    resid = np.ones(n)
    F = np.zeros((n, r))          # rows kept in permuted order while factorising
    k = 0
    pivots = []
    while k < r and resid[k:].sum() > tol * n:
        if k > 0:
            j = k + int(np.argmax(resid[k:]))
            perm[[k, j]] = perm[[j, k]]
            F[[k, j], :k + 1] = F[[j, k], :k + 1]
        else:
            j = 0
        pivots.append(int(perm[k]))
        F[k, k] = math.sqrt(resid[j])   # resid itself is not permuted; position j still holds the pivot's value
        col = np.exp(-omega * (pos[perm[k + 1:]] - pos[perm[k]]) ** 2)
        F[k + 1:, k] = (col - F[k + 1:, :k] @ F[k, :k]) / F[k, k]
        resid[k + 1:] = 1 - np.square(F[k + 1:, :k + 1]).sum(axis=1)
        k += 1
    G = F[np.argsort(perm), :]
    if return_pivots:
        return G, np.asarray(pivots, dtype=np.int64)
    return G


def make_cholesky(lengths, omega, sigma, rank):
    """{length: (L, length, rank)} prior factors, one per unique length -- vlgp/gp.py:150-162."""
    out = {}
    for t in np.unique(np.asarray(lengths)):
        out[int(t)] = np.array([ichol_gauss(int(t), omega[l], rank) * sigma[l] for l in range(len(omega))])
    return out


# --------------------------------------------------------------------------------------------------------------------
# E-step
# --------------------------------------------------------------------------------------------------------------------
def _linear_predictor(x, mu, a, b):
    """eta = mu a + sum_j x[:, j, :] b[j, :] -- vlgp/core.py:66,69."""
    return mu @ a + np.einsum("tjn,jn->tn", x, b)


def _weights(eta, v, a, noise, poisson):
    """w = U (a^T)^2 with U = rate (Poisson) or 1/noise (Gaussian) -- vlgp/core.py:100-104, 437-442."""
    rate = trunc_exp(eta + 0.5 * (v @ (a * a)))
    U = np.where(poisson[None, :], rate, 1.0 / noise[None, :])
    return U @ (a.T ** 2)


def _variance_column(G, w_l):
    """Posterior marginal variance of one latent: rowsum(G o (G - G A + G A (I+A)^-1 A)), A = G' diag(w) G
    -- vlgp/core.py:107-111, 459-468.  Raises LinAlgError if I + A is not positive definite."""
    A = G.T @ (w_l[:, None] * G)
    M = _spd_solve(np.eye(A.shape[0]) + A, A)
    return np.sum(G * (G - G @ A + G @ (A @ M)), axis=1)


def estep_trial(y, x, mu, v, w, a, b, noise, poisson, G, n_iter, dmu_bound=5.0, method="VB"):
    """Variational E-step of ONE trial/segment; returns new (mu, v, w, dmu).  Restates vlgp/core.py:22-120.

    Per iteration: (1) rate from the current q; (2) for every latent, a Newton step on the posterior mean using the
    rank-r prior factor and the *previous* weights (Jacobi over latents: the rate is not refreshed inside the latent
    loop, core.py:76-97); clip to +-dmu_bound; (3) recompute the rate, weights w; (4) if method == "VB" refresh the
    marginal variances v.  A failed r x r solve zeroes that latent's step (core.py:92-94) / skips its v (core.py:112).
    """
    y = np.asarray(y, dtype=float)
    mu = np.array(mu, dtype=float)
    v = np.array(v, dtype=float)
    w = np.array(w, dtype=float)
    dmu = np.zeros_like(mu)
    L = mu.shape[1]
    rank = G.shape[-1]
    eye = np.eye(rank)
    xb = np.einsum("tjn,jn->tn", x, b)
    for _ in range(n_iter):
        eta = mu @ a + xb
        rate = trunc_exp(eta + 0.5 * (v @ (a * a)))
        resid = np.where(poisson[None, :], y - rate, (y - eta) / noise[None, :])
        for l in range(L):
            Gl = G[l]
            wG = w[:, [l]] * Gl
            A = Gl.T @ wG
            u = Gl @ (Gl.T @ (resid @ a[l])) - mu[:, l]
            try:
                c = wG.T @ u
                m = _spd_solve(eye + A, c)
                step = u - Gl @ c + Gl @ (A @ m)
                _clip_inplace(step, dmu_bound)
            except (LinAlgError, ValueError):
                step = np.zeros(mu.shape[0])
            dmu[:, l] = step
            mu[:, l] += step
        eta = mu @ a + xb
        w = _weights(eta, v, a, noise, poisson)
        if method == "VB":
            for l in range(L):
                try:
                    v[:, l] = _variance_column(G[l], w[:, l])
                except (LinAlgError, ValueError):
                    pass
    return mu, v, w, dmu


def estep(trials, params, config, n_iter=None):
    """E-step over a list of trial dicts, in place -- vlgp/core.py:123-126.  ``params['cholesky']`` must hold the factor
    for every trial length."""
    n_iter = config["Eniter"] if n_iter is None else n_iter
    if n_iter < 1:
        return
    poisson = np.asarray(params["likelihood"]) == "poisson"
    for tr in trials:
        G = params["cholesky"][tr["y"].shape[0]]
        mu, v, w, dmu = estep_trial(tr["y"], tr["x"], tr["mu"], tr["v"], tr["w"], params["a"], params["b"],
                                    params["noise"], poisson, G, n_iter, config["dmu_bound"], config["method"])
        tr["mu"][...] = mu
        tr["v"][...] = v
        tr["w"] = w
        tr["dmu"] = dmu


def update_w(trials, params, config=None):
    """w for every trial from the current (mu, v, a, b) -- vlgp/core.py:419-442."""
    poisson = np.asarray(params["likelihood"]) == "poisson"
    for tr in trials:
        tr.setdefault("w", np.zeros_like(tr["mu"]))
        tr.setdefault("v", np.zeros_like(tr["mu"]))
        eta = _linear_predictor(tr["x"], tr["mu"], params["a"], params["b"])
        tr["w"] = _weights(eta, tr["v"], params["a"], params["noise"], poisson)


def update_v(trials, params, config):
    """v for every trial from the current w -- vlgp/core.py:445-471."""
    if config["method"] != "VB":
        return
    for tr in trials:
        tr.setdefault("w", np.zeros_like(tr["mu"]))
        tr.setdefault("v", np.zeros_like(tr["mu"]))
        G = params["cholesky"][tr["mu"].shape[0]]
        for l in range(params["zdim"]):
            try:
                tr["v"][:, l] = _variance_column(G[l], tr["w"][:, l])
            except LinAlgError:
                pass


# --------------------------------------------------------------------------------------------------------------------
# M-step
# --------------------------------------------------------------------------------------------------------------------
def mstep_arrays(y, x, mu, v, a, b, poisson, n_iter, use_hessian=True, eps=1e-8, learning_rate=1.0,
                 da_bound=5.0, db_bound=5.0):
    """M-step on concatenated bins; returns (a, b, noise, da, db).  Restates vlgp/core.py:173-244.

    Per iteration the rate is computed ONCE (core.py:174-176) and every neuron is updated from it: Poisson neurons take
    a clipped Newton step on the loading column then on the regression column (core.py:180-220); Gaussian neurons get
    the closed-form least-squares loading and bias (core.py:221-235).  noise = var(y - eta) per neuron (core.py:177).
    """
    a = np.array(a, dtype=float)
    b = np.array(b, dtype=float)
    y = np.asarray(y, dtype=float)
    L, N = a.shape
    da = np.zeros_like(a)
    db = np.zeros_like(b)
    noise = None
    for _ in range(n_iter):
        eta = mu @ a + np.einsum("tjn,jn->tn", x, b)
        rate = trunc_exp(eta + 0.5 * (v @ (a * a)))
        noise = np.var(y - eta, axis=0, ddof=0)
        for n in range(N):
            xn = x[..., n]
            if poisson[n]:
                shifted = mu + v * a[:, n]
                g = mu.T @ y[:, n] - shifted.T @ rate[:, n]
                if use_hessian:
                    H = shifted.T @ (rate[:, [n]] * shifted)
                    H[np.diag_indices(L)] += rate[:, n] @ v
                    try:
                        step = _spd_solve(H + eps * np.eye(L), g)
                    except (LinAlgError, ValueError):
                        step = learning_rate * g
                else:
                    step = learning_rate * g
                _clip_inplace(step, da_bound)
                da[:, n] = step
                a[:, n] += step

                gb = xn.T @ (y[:, n] - rate[:, n])
                if use_hessian:
                    Hb = xn.T @ (rate[:, [n]] * xn)
                    try:
                        stepb = _spd_solve(Hb + eps * np.eye(Hb.shape[0]), gb)
                    except (LinAlgError, ValueError):
                        stepb = learning_rate * gb
                else:
                    stepb = learning_rate * gb
                _clip_inplace(stepb, db_bound)
                db[:, n] = stepb
                b[:, n] += stepb
            else:
                M = mu.T @ mu
                M[np.diag_indices(L)] += v.sum(axis=0)
                a[:, n] = _spd_solve(M, mu.T @ (y[:, n] - xn @ b[:, n]))
                b[:, n] = _spd_solve(xn.T @ xn, xn.T @ (y[:, n] - mu @ a[:, n]))
                b[1:, n] = 0
    return a, b, noise, da, db


def mstep(trials, params, config):
    """M-step over trial dicts, in place on params -- vlgp/core.py:129-249."""
    if config["Mniter"] < 1:
        return
    y = np.concatenate([tr["y"] for tr in trials], axis=0)
    x = np.concatenate([tr["x"] for tr in trials], axis=0)
    mu = np.concatenate([tr["mu"] for tr in trials], axis=0)
    v = np.concatenate([tr["v"] for tr in trials], axis=0)
    poisson = np.asarray(params["likelihood"]) == "poisson"
    a, b, noise, da, db = mstep_arrays(y, x, mu, v, params["a"], params["b"], poisson, config["Mniter"],
                                       config["use_hessian"], config["eps"], config["learning_rate"],
                                       config["da_bound"], config["db_bound"])
    # the reference works on the arrays stored in params (a, b, da, db are updated IN PLACE, vlgp/core.py:148-156,201,
    # 219; noise is rebound, :177,244) -- observable through FactorAnalysis.transform, whose components_ IS params["a"]
    for key, val in (("a", a), ("b", b), ("da", da), ("db", db)):
        cur = params.get(key)
        if isinstance(cur, np.ndarray) and cur.shape == val.shape and cur.dtype == val.dtype and cur.flags.writeable:
            cur[...] = val
        else:
            params[key] = val
    params["noise"] = noise


# --------------------------------------------------------------------------------------------------------------------
# H-step (GP hyperparameters)
# --------------------------------------------------------------------------------------------------------------------
def se_kernel(t, hyper):
    """K = sigma^2 exp(-omega D^2) + eps I and its derivative w.r.t. log omega -- vlgp/gp.py:46-62.
    (Only the log-omega slot of the reference's dK stack survives its mask [0,1,0], gp.py:16,85.)"""
    sigmasq, omega, eps = hyper
    D2 = (t[:, None] - t[None, :]) ** 2
    Ks = sigmasq * np.exp(-omega * D2)
    dK = -Ks * D2 * omega
    K = Ks + eps * np.eye(len(t))
    return K, dK


def posterior_cov(t, w, hyper):
    """S_i = (K^-1 + diag(w_i))^-1 for every column w_i of w (W x S) -- vlgp/gp.py:126-147.
    ``hyper`` is modified in place if K is not PD (the reference adds log 10 to omega and retries, gp.py:133-135)."""
    while True:
        K, _ = se_kernel(t, hyper)
        try:
            cK = cho_factor(K, lower=True)
            break
        except LinAlgError:
            hyper[1] += np.log(10)
    n = len(t)
    Kinv = cho_solve(cK, np.eye(n))
    S = np.empty((n, n, w.shape[1]))
    for i in range(w.shape[1]):
        S[:, :, i] = cho_solve(cho_factor(Kinv + np.diag(w[:, i]), lower=True), np.eye(n))
    return S


def elbo(hyper, t, mu, S):
    """(ll, dll/dlog omega) of vlgp/gp.py:12-43 with mask [0,1,0]; mu is (W x S), S is (W x W x S).
    Returns (-inf, 0.0) if K is not PD (gp.py:17-20)."""
    K, dK = se_kernel(t, hyper)
    try:
        cK = cho_factor(K, lower=True)
    except LinAlgError:
        return -np.inf, 0.0
    n = len(t)
    Kinv = cho_solve(cK, np.eye(n))
    alpha = cho_solve(cK, mu)
    nseg = mu.shape[1]
    ll = -0.5 * np.einsum("ik,ik->", mu, alpha)
    acc = alpha @ alpha.T - nseg * Kinv
    for i in range(nseg):
        KiS = cho_solve(cK, S[:, :, i])
        ll -= 0.5 * np.trace(KiS)
        acc += KiS @ Kinv
    ll -= nseg * np.log(np.diag(cK[0])).sum()
    dll = 0.5 * np.sum(acc * dK)
    return ll, dll


def hstep_objective(logp, t, mu, w):
    """-(ll), -(dll) as a 3-vector in log(sigma^2, omega, eps) -- the closure at vlgp/gp.py:107-111."""
    hyper = np.exp(logp)
    S = posterior_cov(t, w, hyper)
    ll, dll = elbo(hyper, t, mu, S)
    return -ll, -np.array([0.0, dll, 0.0])


def hstep(trials, params, config):
    """Per-latent L-BFGS-B over log(sigma^2, omega, eps); only omega moves -- vlgp/core.py:252-257, vlgp/gp.py:65-123."""
    if not config["Hstep"]:
        return
    from scipy.optimize import minimize

    mu = np.stack([tr["mu"] for tr in trials])
    w = np.stack([tr["w"] for tr in trials])
    t = np.arange(config["window"]) * params["dt"]
    gp_noise = params["gp_noise"]
    sigma, omega = params["sigma"], params["omega"]
    n_eval = []
    for l in range(params["zdim"]):
        x0 = np.log((sigma[l] ** 2, omega[l], gp_noise))
        bounds = np.log(((1e-3, 1), config["omega_bound"], (gp_noise / 2, gp_noise * 2)))
        res = minimize(hstep_objective, x0, args=(t, mu[:, :, l].T, w[:, :, l].T), jac=True, bounds=bounds)
        sigmasq, omega_new, _ = np.exp(res.x)
        if not np.any(np.isclose(omega_new, config["omega_bound"])):
            omega[l] = omega_new
        sigma[l] = np.sqrt(sigmasq)
        n_eval.append(res.nfev)
    params["sigma"], params["omega"] = sigma, omega
    params["cholesky"] = make_cholesky([tr["y"].shape[0] for tr in trials], omega, sigma, params["rank"])
    return n_eval


# --------------------------------------------------------------------------------------------------------------------
# constraints, outer loop
# --------------------------------------------------------------------------------------------------------------------
def constrain_loading(trials, params, config):
    """'fro' (default) / row-norm / 'svd' normalisation of the loading, compensated in mu -- vlgp/core.py:392-416."""
    kind = config["constrain_loading"]
    if not kind or kind == "none":
        return
    a = params["a"]
    if kind == "svd":
        _, _, vt = np.linalg.svd(a, full_matrices=False)
        us = a @ vt.T
        for tr in trials:
            tr["mu"] = tr["mu"] @ us
        params["a"] = vt
        return
    if kind == "fro":
        s = np.linalg.norm(a) + config["eps"]
        a /= s
        for tr in trials:
            tr["mu"] *= s
    else:
        s = np.linalg.norm(a, ord=kind, axis=1, keepdims=True) + config["eps"]
        a /= s
        for tr in trials:
            tr["mu"] *= s.T


def constrain_latent(trials, params, config):
    """Optional centring / scaling of mu, compensated in b / a -- vlgp/core.py:366-389."""
    kind = config["constrain_latent"]
    if not kind or kind == "none":
        return
    mu = np.concatenate([tr["mu"] for tr in trials], axis=0)
    mean = mu.mean(axis=0, keepdims=True)
    std = mu.std(axis=0, keepdims=True)
    if kind in ("location", "both"):
        for tr in trials:
            tr["mu"] -= mean
        params["b"][0, :] += np.squeeze(mean @ params["a"])
    if kind in ("scale", "both"):
        for tr in trials:
            tr["mu"] /= std
        params["a"] *= std.T


def default_config(**kw):
    """Defaults of vlgp/preprocess.py:84-112 (unknown keys dropped)."""
    cfg = {
        "constrain_loading": "fro", "constrain_latent": False, "use_hessian": True, "eps": 1e-8, "tol": 1e-8,
        "min_iter": 5, "method": "VB", "learning_rate": 1.0, "max_iter": 20, "Eniter": 25, "Mniter": 25,
        "Hstep": True, "da_bound": 5.0, "db_bound": 5.0, "dmu_bound": 5.0, "omega_bound": (5e-4, 5e-2),
        "window": 50, "saving_interval": 1800, "callbacks": [], "parallel": False,
    }
    cfg.update({k: v for k, v in kw.items() if k in cfg})
    return cfg


def vem(trials, params, config):
    """Outer variational-EM loop on (already cut) segments, in place -- vlgp/core.py:269-359.
    Fills config['runtime'] with the same keys as the reference."""
    rt = {"it": 0, "e_elapsed": [], "m_elapsed": [], "h_elapsed": [], "em_elapsed": []}
    tol = config["tol"]
    for it in range(config["max_iter"]):
        rt["it"] += 1
        n_mu = np.linalg.norm(np.concatenate([tr["mu"] for tr in trials], axis=0))
        n_a = np.linalg.norm(params["a"])
        n_b = np.linalg.norm(params["b"])
        t0 = time.perf_counter()
        constrain_loading(trials, params, config)
        estep(trials, params, config)
        t1 = time.perf_counter()
        constrain_latent(trials, params, config)
        mstep(trials, params, config)
        t2 = time.perf_counter()
        hstep(trials, params, config)
        t3 = time.perf_counter()
        rt["e_elapsed"].append(t1 - t0)
        rt["m_elapsed"].append(t2 - t1)
        rt["h_elapsed"].append(t3 - t2)
        rt["em_elapsed"].append(t3 - t0)
        config["runtime"] = rt
        for cb in config["callbacks"]:
            try:
                cb(trials, params, config)
            except RuntimeError:
                pass
        n_dmu = np.linalg.norm(np.concatenate([tr["dmu"] for tr in trials], axis=0))
        done = (n_dmu < tol * n_mu and np.linalg.norm(params["da"]) < tol * n_a
                and np.linalg.norm(params["db"]) < tol * n_b)
        if done and it + 1 >= config["min_iter"]:
            break


def cut_trial_starts(length, window, rng_multinomial=np.random.multinomial):
    """Start bins of the ceil(length/window) windows of one trial -- vlgp/util.py:482-494.  Draws ONE multinomial from
    the global NumPy RNG per trial (also when the overlap is zero), like the reference."""
    nseg = math.ceil(length / window)
    overlap = nseg * window - length
    start = np.cumsum(np.full(nseg, window, dtype=int)) - window
    shift = np.cumsum(np.append([0], rng_multinomial(overlap, np.ones(nseg - 1) / (nseg - 1))))
    return start - shift


# --------------------------------------------------------------------------------------------------------------------
# GPFA branch (vlgp/gpfa.py:20-56)
# --------------------------------------------------------------------------------------------------------------------
def sekernel(x, var, scale, jitter=1e-6):
    """vlgp/gp.py:165-171."""
    x = np.asarray(x, dtype=float).reshape(-1, 1) / scale
    return var * np.exp(-0.5 * (x - x.T) ** 2) + np.eye(x.shape[0]) * jitter


def gpfa_em(y, C, d, R, K, max_iter):
    """Restates gpfa.em (vlgp/gpfa.py:20-56) without its (n ydim)^2 Kronecker matrices: E-step through the push-through
    identity z = (I + bigK S)^-1 bigK bigC' bigR^-1 (y - d), M-step lstsq through its normal equations.  Keeps the
    reference's two quirks: bigR is built once from the R passed in (:31), and it is ordered (time, neuron) against the
    (neuron, time) order of bigC and the residual (:30,41-42), i.e. noise R[(j n + t) % ydim] for neuron j at bin t."""
    y = np.asarray(y, dtype=float)
    m, n, ydim = y.shape
    C = np.array(C, dtype=float)
    zdim = C.shape[0]
    d = np.asarray(d, dtype=float).reshape(1, ydim)
    Y = y.reshape(-1, ydim)
    bigK = np.kron(np.eye(zdim), K)
    idx = (np.arange(ydim)[:, None] * n + np.arange(n)[None, :]) % ydim
    rho = 1.0 / np.diag(R)[idx]                                        # [j, t]
    z = np.zeros((m, n, zdim))
    t = np.arange(n)
    for _ in range(max_iter):
        h = np.einsum("lj,jt,stj->slt", C, rho, y - d[None, :])
        Q = np.einsum("lj,kj,jt->lkt", C, C, rho)
        S = np.zeros((zdim * n, zdim * n))
        for l in range(zdim):
            for k in range(zdim):
                S[l * n + t, k * n + t] = Q[l, k]
        P = np.linalg.solve(np.eye(zdim * n) + bigK @ S, bigK)
        z = (h.reshape(m, -1) @ P.T).reshape(m, zdim, n).transpose(0, 2, 1)
        z = z - z.mean(axis=(0, 1), keepdims=True)
        Z1 = np.column_stack([z.reshape(-1, zdim), np.ones(m * n)])
        coef, r, *_ = np.linalg.lstsq(Z1, Y, rcond=None)
        C, d = coef[:-1, :], coef[[-1], :]
        R = np.diag(r ** 2)
        C = C / np.linalg.norm(C)
    return z, C, d, R
