"""TEST INFRASTRUCTURE ONLY: a stand-in for vlgp_b200.engine.Engine / TrialSet whose "device" is the NumPy oracle.

It lets the CPU suite drive the package's real HOST code -- api.fit / transform, core.vem and its Session, the
lockstep L-BFGS-B driver of gp.py, the in-place / rebinding rules of the trial and params dicts -- end to end and compare
the result with the golden vectors of the unmodified reference, without a GPU.  Nothing under vlgp_b200/ imports this
file; the product has no CPU path (tests/test_abi.py::test_no_cpu_fallback)."""
from __future__ import annotations

import numpy as np

from oracle import vlgp_oracle as orc


class OracleEngine:
    """Mirrors the attributes and methods of vlgp_b200.engine.Engine that the host code uses."""

    def __init__(self):
        self.world_size, self.rank_id, self.device = 1, 0, 0
        self.N = self.L = self.rank = 0
        self.params = {}
        self.model_key = None
        self.log = []               # (method name, detail) in call order, for orchestration checks

    # -- model / parameters ------------------------------------------------------------------------------------------
    def ensure_model(self, params):
        lik = np.asarray(params["likelihood"])
        if (~np.isin(lik, ("poisson", "gaussian"))).any():
            raise ValueError("unsupported likelihood")
        if int(params.get("xdim", 1)) != 1:
            raise NotImplementedError("xdim == 1 only")
        self.N, self.L, self.rank = int(params["ydim"]), int(params["zdim"]), int(params["rank"])
        self.poisson = lik == "poisson"
        self.gp_noise, self.dt = float(params["gp_noise"]), float(params["dt"])

    def push_params(self, params, which=("a", "b", "noise", "sigma", "omega")):
        self.log.append(("push_params", tuple(which)))
        for k in which:
            val = np.array(params[k], dtype=float)
            self.params[k] = val.reshape(1, self.N) if k == "b" else val

    def pull_params(self, params, which=("a", "b", "noise", "da", "db")):
        from vlgp_b200.util import assign_inplace

        self.log.append(("pull_params", tuple(which)))
        for k in which:
            val = np.array(self.params[k], dtype=float)
            if k == "noise":
                params[k] = val
            else:
                assign_inplace(params, k, val)
        return params

    def new_trials(self, lengths):
        return OracleTrialSet(self, lengths)

    def allreduce(self, x, op="sum"):
        return np.asarray(x, dtype=float)

    def sync(self):
        pass

    # -- measurement entry points (bench.py): fixed stand-in values, only the FLOW is under test -----------------------
    def peak_fp64(self):
        return {"dfma_tflops": 30.0, "dmma_tflops": 35.0}

    def peak_hbm(self, nbytes=1 << 30):
        return 6000.0

    def flush_l2(self):
        pass

    def set_precision(self, bits=64):
        self.precision = int(bits)        # the stand-in always computes in double precision

    def timer_start(self):
        import time

        self._t0 = time.perf_counter()

    def timer_stop(self):
        import time

        return (time.perf_counter() - self._t0) * 1e3

    def counters(self):
        return {"launches": len(self.log), "estep_solves": 0, "hstep_evals": 0, "allreduces": 0}

    def profile_enable(self, mask=0xF):
        self._prof = mask

    def profile_get(self, which):
        return 1.0, 1

    def oracle_params(self):
        p = dict(self.params)
        p.update(zdim=self.L, ydim=self.N, xdim=1, rank=self.rank, gp_noise=self.gp_noise, dt=self.dt,
                 likelihood=np.where(self.poisson, "poisson", "gaussian"))
        return p


class OracleTrialSet:
    def __init__(self, eng, lengths):
        self.eng = eng
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        self.nbin = int(self.lengths.sum())
        self.starts = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)
        L, N = eng.L, eng.N
        self.y = np.zeros((self.nbin, N))
        self.state = {k: np.zeros((self.nbin, L)) for k in ("mu", "v", "w", "dmu")}
        self.chol = {}
        self.h2d_bytes = self.d2h_bytes = 0
        self.id = 0
        self._pending = None

    def free(self):
        self.id = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()
        return False

    # -- data ----------------------------------------------------------------------------------------------------------
    def _split(self, arr):
        return [arr[s:s + n] for s, n in zip(self.starts, self.lengths)]

    def _trials(self):
        """Views of the resident arrays as the oracle's list of trial dicts (writes go through)."""
        N = self.eng.N
        out = []
        for i, (s, n) in enumerate(zip(self.starts, self.lengths)):
            tr = {"y": self.y[s:s + n], "x": np.ones((n, 1, N))}
            for k in ("mu", "v", "w", "dmu"):
                tr[k] = self.state[k][s:s + n]
            out.append(tr)
        return out

    def _store(self, trials, keys):
        for tr, s, n in zip(trials, self.starts, self.lengths):
            for k in keys:
                self.state[k][s:s + n] = tr[k]

    def set_y_parts(self, ys):
        if len(ys) != self.lengths.size:
            raise ValueError("expected %d observation blocks" % self.lengths.size)
        for blk, dst in zip(ys, self._split(self.y)):
            dst[...] = np.asarray(blk, dtype=float)
        return 0

    def set_y(self, y, ydtype=None):
        self.y[...] = np.asarray(y, dtype=float)

    def project_y(self, mean, P, Cz):
        """FactorAnalysis.transform of every bin (vlgp/preprocess.py:36-41) -- the engine's device projection, here with
        scikit-learn's own order of operations per trial."""
        for blk, dst in zip(self._split(self.y), self._split(self.state["mu"])):
            dst[...] = np.dot(np.dot(blk - np.asarray(mean, dtype=float), np.asarray(P, dtype=float)), Cz)

    def set_state_parts(self, **blocks):
        for key, arrs in blocks.items():
            if arrs is None:
                continue
            for blk, dst in zip(arrs, self._split(self.state[key])):
                dst[...] = blk

    def get_state_parts(self, **blocks):
        for key, arrs in blocks.items():
            if arrs is None:
                continue
            if len(arrs) == 1 and len(self.lengths) > 1:          # one block for the whole set
                if arrs[0].shape != self.state[key].shape or not arrs[0].flags.writeable:
                    raise ValueError("block shape")
                arrs[0][...] = self.state[key]
                continue
            for blk, src in zip(arrs, self._split(self.state[key])):
                if not isinstance(blk, np.ndarray) or blk.dtype != np.float64 or blk.shape != src.shape \
                        or not blk.flags.writeable or not blk.flags.c_contiguous:
                    raise ValueError("block is not a writable C-contiguous float64 array of the trial's shape")
                blk[...] = src

    def set_state(self, mu=None, v=None, w=None):
        for k, val in (("mu", mu), ("v", v), ("w", w)):
            if val is not None:
                self.state[k][...] = val

    def get_state(self, which=("mu", "v", "w", "dmu")):
        return {k: self.state[k].copy() for k in which}

    # -- prior factor -------------------------------------------------------------------------------------------------
    def make_cholesky(self):
        p = self.eng.params
        self.eng.log.append(("make_cholesky", tuple(sorted(set(self.lengths.tolist())))))
        self.chol = orc.make_cholesky(sorted(set(self.lengths.tolist())), p["omega"], p["sigma"], self.eng.rank)

    def get_cholesky(self, length, with_pivots=False):
        return self.chol[int(length)].copy()

    def set_cholesky(self, length, G):
        self.chol[int(length)] = np.array(G, dtype=float)

    def _params(self):
        p = self.eng.oracle_params()
        p["cholesky"] = self.chol
        return p

    # -- steps -----------------------------------------------------------------------------------------------------------
    row_ops = True      # E-step over a subset of segments, row copies, affine map on selected rows (see core.py)

    def estep(self, n_iter, dmu_bound=5.0, method="VB", subset=None):
        self.eng.log.append(("estep", n_iter if subset is None else (n_iter, len(subset))))
        trials = self._trials()
        sel = range(len(trials)) if subset is None else [int(i) for i in subset]
        work = [dict(trials[i], **{k: trials[i][k].copy() for k in ("mu", "v", "w", "dmu")}) for i in sel]
        cfg = orc.default_config(dmu_bound=dmu_bound, method=method, Eniter=n_iter)
        orc.estep(work, self._params(), cfg, n_iter=n_iter)
        for i, tr in zip(sel, work):
            s, n = self.starts[i], self.lengths[i]
            for k in ("mu", "v", "w", "dmu"):
                self.state[k][s:s + n] = tr[k]
        return 0

    def copy_rows(self, src, dst, which=("mu", "v")):
        src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
        for k in which:
            self.state[k][dst] = self.state[k][src]

    def update_w(self):
        self.eng.log.append(("update_w", None))
        trials = self._trials()
        orc.update_w(trials, self._params())
        self._store(trials, ("w",))

    def update_v(self):
        self.eng.log.append(("update_v", None))
        trials = self._trials()
        orc.update_v(trials, self._params(), orc.default_config())
        self._store(trials, ("v",))
        return 0

    def _mstep(self, n_iter, use_hessian, eps, learning_rate, da_bound, db_bound):
        p = self._params()
        p["da"], p["db"] = np.zeros_like(p["a"]), np.zeros_like(p["b"])
        cfg = orc.default_config(Mniter=n_iter, use_hessian=use_hessian, eps=eps, learning_rate=learning_rate,
                                 da_bound=da_bound, db_bound=db_bound)
        orc.mstep(self._trials(), p, cfg)
        for k in ("a", "b", "noise", "da", "db"):
            self.eng.params[k] = np.array(p[k], dtype=float)

    def mstep(self, n_iter, use_hessian=True, eps=1e-8, learning_rate=1.0, da_bound=5.0, db_bound=5.0):
        self.eng.log.append(("mstep", n_iter))
        self._mstep(n_iter, use_hessian, eps, learning_rate, da_bound, db_bound)
        return 0

    def mstep_begin(self, n_iter, use_hessian=True, eps=1e-8, learning_rate=1.0, da_bound=5.0, db_bound=5.0):
        # the device runs it concurrently with the H-step; both read only what the E-step left, so running it at
        # _end (after the H-step) on a snapshot of mu / v taken now gives the same numbers and checks exactly that
        # independence
        assert self._pending is None, "mstep_begin twice"
        self.eng.log.append(("mstep_begin", n_iter))
        self._pending = (n_iter, use_hessian, eps, learning_rate, da_bound, db_bound, self.state["mu"].copy(),
                         self.state["v"].copy(), {k: np.array(self.eng.params[k]) for k in ("a", "b", "noise")})

    def mstep_end(self):
        if self._pending is None:
            return 0
        *args, mu, v, pab = self._pending
        self._pending = None
        self.eng.log.append(("mstep_end", None))
        assert np.array_equal(mu, self.state["mu"]) and np.array_equal(v, self.state["v"]), \
            "the H-step changed mu / v while the M-step was pending"
        for k in ("a", "b", "noise"):
            assert np.array_equal(pab[k], self.eng.params[k]), "a / b / noise changed while the M-step was pending"
        self._mstep(*args)
        return 0

    def hstep_prepare(self):
        self.eng.log.append(("hstep_prepare", None))
        W = int(self.lengths[0])
        assert (self.lengths == W).all()
        S = len(self.lengths)
        self._h_mu = self.state["mu"].reshape(S, W, self.eng.L)
        self._h_w = self.state["w"].reshape(S, W, self.eng.L)
        self._h_t = np.arange(W) * self.eng.dt

    def hstep_objective(self, latent, hyper):
        hyper = np.array(hyper, dtype=float)
        K, _ = orc.se_kernel(self._h_t, hyper)
        try:
            np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            return 0.0, 0.0, 1
        S = orc.posterior_cov(self._h_t, self._h_w[:, :, latent].T, hyper)
        ll, dll = orc.elbo(hyper, self._h_t, self._h_mu[:, :, latent].T, S)
        return float(ll), float(dll), 0

    def hstep_objective_batch(self, latents, hypers):
        self.eng.log.append(("hstep_round", len(latents)))
        out = [self.hstep_objective(int(l), h) for l, h in zip(latents, np.asarray(hypers, dtype=float))]
        return (np.array([o[0] for o in out]), np.array([o[1] for o in out]), np.array([o[2] for o in out], np.int32))

    def latent_affine(self, shift=None, M=None, rows=None):
        mu = self.state["mu"]
        idx = slice(None) if rows is None else np.asarray(rows, dtype=np.int64)
        x = mu[idx]
        if shift is not None:
            x = x - np.asarray(shift, dtype=float).reshape(1, -1)
        if M is not None:
            x = x @ np.asarray(M, dtype=float)
        mu[idx] = x

    def norms(self):
        return float(np.sum(self.state["mu"] ** 2)), float(np.sum(self.state["dmu"] ** 2))

    def latent_moments(self):
        mu = self.state["mu"]
        return mu.sum(axis=0), (mu ** 2).sum(axis=0), int(self.nbin)


def install(monkeypatch):
    """Make vlgp_b200.engine.get_engine() return a fresh OracleEngine for the duration of a test."""
    import vlgp_b200.engine as engine_mod

    eng = OracleEngine()
    monkeypatch.setattr(engine_mod, "_ENGINE", eng)
    return eng
