"""Variational-EM engine: the reference's ``(trials, params, config)`` in-place functions, executed on the GPU.

Drop-in for vlgp/core.py: ``vem`` (:269-359), ``estep`` (:123-126), ``mstep`` (:129-249), ``hstep`` (:252-257),
``infer`` (:260-266), ``update_w`` (:419-442), ``update_v`` (:445-471), ``constrain_loading`` (:392-416),
``constrain_latent`` (:366-389).  Each function packs the dicts into a device ``Session``, calls the C ABI
(include/vlgp_b200.h) and unpacks the result with the reference's aliasing rules: ``mu`` and ``v`` are updated IN
PLACE (segments made by ``cut_trials`` are views of their trial), ``w`` and ``dmu`` are rebound.  ``vem`` keeps one
session for the whole loop, so host<->device traffic is one upload before and one download after the iterations
(plus the tiny parameter arrays every iteration).  There is no host implementation of any step: without the native
library every function raises ``VlgpNativeError``.
"""
from __future__ import annotations

import logging
import os
import time
import weakref

import numpy as np

from .engine import Engine, TrialSet, get_engine, _fastpack
from .util import assign_inplace

__all__ = ["vem", "estep", "mstep", "hstep", "infer", "update_w", "update_v", "constrain_loading", "constrain_latent",
           "Session"]

logger = logging.getLogger(__name__)


def _echo(msg):
    try:
        import click

        click.echo(msg)
    except Exception:  # pragma: no cover
        print(msg)


_ONES_VERIFIED = {}      # id(array) -> weakref: regressor arrays already known to be all ones


def _all_ones(x):
    """True if the (bins, 1, neurons) regressor is the all-ones bias column.  Segments made by cut_trials are views of
    their trial's x: the trial-level array is scanned once and remembered (weakly), not once per segment per call."""
    base = x
    while isinstance(base.base, np.ndarray):
        base = base.base
    ref = _ONES_VERIFIED.get(id(base))
    if ref is not None and ref() is base:
        return True
    if base.min() != 1 or base.max() != 1:
        return base is not x and x.min() == 1 and x.max() == 1
    _remember_ones(base)
    return True


def _remember_ones(base):
    """Record (weakly) that ``base`` -- an array that owns its memory -- is all ones."""
    try:
        key = id(base)
        _ONES_VERIFIED[key] = weakref.ref(base, lambda _r, k=key: _ONES_VERIFIED.pop(k, None))
    except TypeError:  # pragma: no cover
        pass


def _bias_only(trials, params):
    """True when every trial's regressor x is the all-ones bias column with xdim == 1 (the reference's default,
    vlgp/preprocess.py:43-44): the tuned kernels then treat the regression as a per-neuron constant b[n].  Anything else
    -- xdim = max(history, 1) > 1 or a user-supplied design -- is uploaded and handled by csrc/regress.cu.
    Segments made by cut_trials are views of their trial's x, so the scan runs once per underlying array (and at most
    once per process per array, see _all_ones): 5120 segments of 256 trials cost 256 scans, and finding those 256 is
    one pass in C (_fastpack)."""
    if int(params.get("xdim", 1) or 1) != 1:
        return False
    xs = [x for x in (tr.get("x") for tr in trials) if x is not None]     # a missing x is the default regressor
    if not xs:
        return True
    owners = None
    if _fastpack is not None:
        try:
            owners = _fastpack.block_owners(xs, 3, 1, 1)
        except (ValueError, TypeError, BufferError, AttributeError):
            owners = None
    if owners is not None and all(o.ndim == 3 and o.shape[1] == 1 for o in owners) and all(_all_ones(o) for o in owners):
        return True
    # views that cover only part of an owner which is not all ones as a whole: judge every view by its own entries
    return all(x.ndim == 3 and x.shape[1] == 1 and _all_ones(x) for x in xs)


def _row_views(a, starts, lengths):
    """Per-trial views of the rows of one (bins, L) block (what np.split returns, without its per-piece overhead)."""
    if lengths.size and (lengths == lengths[0]).all():
        return list(a.reshape(lengths.size, int(lengths[0]), a.shape[1]))
    return [a[s:s + n] for s, n in zip(starts.tolist(), lengths.tolist())]


def _shared_rows(trials):
    """Junctions between consecutive segments that are overlapping views of one trial (windows cut from a trial whose
    length is not a multiple of the window, vlgp/util.py:482-498): list of (k, n) -- the last n rows of segment k ARE
    the first n rows of segment k + 1 in the reference (one block of memory).  Found from the addresses of the mu
    arrays, so it needs nothing beyond what cut_trials returns.  One vectorised pass over the address table (5120
    segments: 0.3 ms); only candidate pairs are looked at one by one."""
    n = len(trials)
    if n < 2:
        return []
    ptr = rows = None
    if _fastpack is not None:          # address table in C: every mu a C-contiguous float64 (rows, L) block, else fall through
        try:
            mus = [tr["mu"] for tr in trials]
            L = mus[0].shape[1]
            pb, rb = _fastpack.pointers(mus, 8, L, False, "d")
            ptr = np.frombuffer(pb, dtype=np.uint64).astype(np.int64)
            rows = np.frombuffer(rb, dtype=np.int64)
            width = np.full(n, L, dtype=np.int64)
            ok = np.ones(n, dtype=bool)
        except (KeyError, TypeError, ValueError, BufferError, AttributeError, IndexError):
            ptr = None
    if ptr is None:
        ptr = np.zeros(n, dtype=np.int64)
        rows = np.zeros(n, dtype=np.int64)
        width = np.zeros(n, dtype=np.int64)
        ok = np.zeros(n, dtype=bool)
        for k, tr in enumerate(trials):
            a = tr.get("mu")
            if isinstance(a, np.ndarray) and a.ndim == 2 and a.dtype == np.float64 and a.flags.c_contiguous:
                ptr[k] = a.__array_interface__["data"][0]
                rows[k], width[k] = a.shape
                ok[k] = True
    step = width[:-1] * 8
    d = ptr[1:] - ptr[:-1]
    cand = ok[:-1] & ok[1:] & (width[:-1] == width[1:]) & (d > 0) & (step > 0)
    cand &= np.where(step > 0, d % np.maximum(step, 1), 1) == 0
    shared = rows[:-1] - d // np.maximum(step, 1)
    cand &= (shared > 0) & (shared < np.minimum(rows[:-1], rows[1:]))
    out = []
    for k in np.flatnonzero(cand):
        if np.may_share_memory(trials[k]["mu"], trials[k + 1]["mu"]):
            out.append((int(k), int(shared[k])))
    return out


class _Aliasing:
    """Row bookkeeping that reproduces what the reference computes when segments share bins.  Every segment has its
    own copy of its rows on the device; the reference has ONE copy of a shared row, processes the segments in list order
    and updates mu / v in place (vlgp/core.py:96-97,112,123-126).  Hence:
      * E-step: a segment starts from the mu, v its predecessor left on the shared rows -> segments are run level by
        level along each chain of overlapping junctions (level = position in the chain), the predecessor's tail rows
        being copied to the successor's head rows before its level runs; afterwards the successor's values are copied
        back, so both copies hold what the reference's single row holds (the M- and H-step read them);
      * constrain_loading / constrain_latent (vlgp/core.py:384-389,414-416) go through the segment list and rescale /
        shift every segment's mu in place: a shared row gets the operation once per segment that contains it."""

    def __init__(self, junctions, starts, lengths):
        n_seg = len(lengths)
        level = np.zeros(n_seg, dtype=np.int64)
        succ = {k for k, _ in junctions}
        for k, _ in junctions:
            level[k + 1] = level[k] + 1
        self.levels = [np.flatnonzero(level == lv) for lv in range(int(level.max()) + 1)]
        self.fwd = []                  # per level >= 1: (src rows in the predecessors, dst rows in this level's segments)
        tail, head = {}, {}
        for k, n in junctions:
            tail[k] = np.arange(starts[k] + lengths[k] - n, starts[k] + lengths[k])
            head[k + 1] = np.arange(starts[k + 1], starts[k + 1] + n)
        for lv in range(1, len(self.levels)):
            segs = [int(i) for i in self.levels[lv]]
            self.fwd.append((np.concatenate([tail[i - 1] for i in segs]), np.concatenate([head[i] for i in segs])))
        ks = sorted(succ)
        self.back_src = np.concatenate([head[k + 1] for k in ks])
        self.back_dst = np.concatenate([tail[k] for k in ks])
        self.shared = np.concatenate([self.back_src, self.back_dst])


class Session:
    """Device copy of a list of trials plus the current parameters."""

    def __init__(self, trials, params, engine: Engine = None, upload_factors=True):
        self.eng = engine or get_engine()
        eng = self.eng
        eng.ensure_model(params)
        eng.push_params(params)
        self.n = len(trials)
        lengths = [tr["y"].shape[0] for tr in trials]
        bias_only = _bias_only(trials, params)
        self.ts: TrialSet = eng.new_trials(lengths)
        try:
            self.ts.set_y_parts([tr["y"] for tr in trials])
            if not bias_only:
                N, xd = eng.N, eng.xdim
                xs = [np.asarray(tr["x"], dtype=np.float64) if tr.get("x") is not None else np.ones((n, xd, N))
                      for tr, n in zip(trials, lengths)]
                if any(x.shape != (n, xd, N) for x, n in zip(xs, lengths)):
                    raise ValueError("every trial's x must have shape (bins, %d, %d)" % (xd, N))
                self.ts.set_x(np.concatenate(xs, axis=0))
            L = eng.L

            def blocks(key):
                try:
                    out = [tr[key] for tr in trials]
                    if not [1 for b in out if b is None]:
                        return out
                except KeyError:
                    pass
                return [tr[key] if tr.get(key) is not None else np.zeros((n, L)) for tr, n in zip(trials, lengths)]

            self.ts.set_state_parts(mu=blocks("mu"), v=blocks("v"), w=blocks("w"))
            # overlapping windows: exact reference semantics need three row-level device operations; without them the
            # segments are treated as independent copies
            self.alias = None
            if getattr(self.ts, "row_ops", False):
                junctions = _shared_rows(trials)
                if junctions:
                    self.alias = _Aliasing(junctions, self.ts.starts, self.ts.lengths)
            chol = params.get("cholesky") if upload_factors else None
            self.have_factors = False
            if chol:
                uniq = sorted(set(lengths))
                if all(t in chol for t in uniq):
                    for t in uniq:
                        self.ts.set_cholesky(t, chol[t])
                    self.have_factors = True
        except Exception:
            self.ts.free()
            raise

    def refresh(self, trials, params, which=("mu", "v", "w")):
        """Re-send parameters and the given per-bin arrays from the host dicts (after another session changed them)."""
        self.eng.push_params(params)
        self.ts.set_state_parts(**{k: [tr[k] for tr in trials] for k in which})

    def require_factors(self):
        if not self.have_factors:
            raise KeyError("params['cholesky'] lacks the prior factor of a trial length: call make_cholesky first")

    def make_cholesky(self, params):
        """Prior factors for this set's lengths from params['sigma'/'omega'], on device; mirrored into
        params['cholesky'] (REPLACING the dict, like vlgp/gp.py:158)."""
        self.eng.push_params(params, which=("sigma", "omega"))
        self.ts.make_cholesky()
        params["cholesky"] = {int(t): self.ts.get_cholesky(int(t)) for t in sorted(set(self.ts.lengths.tolist()))}
        self.have_factors = True

    def pull(self, trials, which=("mu", "v", "w", "dmu"), rebind=()):
        """Device state -> trial dicts with the reference's aliasing: mu and v are written IN PLACE (segment arrays are
        views of their trial), w and dmu are rebound to fresh arrays (views of one new block per key).  Keys listed in
        ``rebind`` are rebound as well: constrain_loading="svd" REPLACES every trial["mu"] (vlgp/core.py:404-405), so
        from then on the segments' posterior means no longer reach the trials they were cut from."""
        L = self.eng.L
        fresh, landed = {}, {}
        for k in which:
            done = False
            if k in ("mu", "v") and k not in rebind:
                try:      # in place when every trial already holds a writable C-contiguous float64 block of its shape
                    self.ts.get_state_parts(**{k: [tr[k] for tr in trials]})
                    done = True
                except (KeyError, TypeError, ValueError):
                    done = False
            if not done:
                take = getattr(self.ts, "take_prefetched", None)
                got = take(k) if take is not None else None       # the block a prefetch filled IS the new array
                if got is not None:
                    landed[k] = got
                else:
                    fresh[k] = np.empty((self.ts.nbin, L))
        if fresh:
            self.ts.get_state_parts(**{k: [a] for k, a in fresh.items()})
        fresh.update(landed)
        if fresh:
            for k, a in fresh.items():
                views = _row_views(a, self.ts.starts, self.ts.lengths)
                if k in ("mu", "v") and k not in rebind:
                    for tr, val in zip(trials, views):
                        if isinstance(tr.get(k), np.ndarray) and tr[k].shape == val.shape:
                            tr[k][...] = val
                        else:
                            tr[k] = val
                else:
                    for tr, val in zip(trials, views):
                        tr[k] = val

    def close(self):
        self.ts.free()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# ----------------------------------------------------------------------------------------------------------------------
# single steps (stateless: upload, run, download)
# ----------------------------------------------------------------------------------------------------------------------
def _with_session(trials, params, session, **kw):
    """An open session of the caller (state already on the device) or a temporary one for this call."""
    import contextlib

    return contextlib.nullcontext(session) if session is not None else Session(trials, params, **kw)


def _precision_bits(config):
    """32 when the configuration asks for single-precision rate passes (config["dtype"] = "float32"), else 64."""
    return 32 if str(config.get("dtype", "float64")).lower() in ("float32", "f32", "fp32", "single") else 64


def estep(trials, params, config, session=None):
    """Update the variational posterior q (E-step) of every trial."""
    if config["Eniter"] < 1:
        return
    with _with_session(trials, params, session) as s:
        s.require_factors()
        s.eng.set_precision(_precision_bits(config))
        if s.alias is None:
            nfail = s.ts.estep(config["Eniter"], config["dmu_bound"], config["method"])
        else:
            nfail = _estep_aliased(s, config)
        if nfail:
            logger.error("E-step: %d r x r systems were not positive definite (update skipped)", nfail)
        s.pull(trials)


def infer(trials, params, config, session=None):
    """E-step with Eniter := max_iter on the given (uncut) trials -- vlgp/core.py:260-266."""
    niter = config["Eniter"]
    config["Eniter"] = config["max_iter"]
    t0 = time.perf_counter()
    try:
        estep(trials, params, config, session=session)
    finally:
        config["Eniter"] = niter
    _echo("{:.2f}s".format(time.perf_counter() - t0))


def update_w(trials, params, config, session=None):
    with _with_session(trials, params, session, upload_factors=False) as s:
        s.ts.update_w()
        s.pull(trials, ("w",))
    for tr in trials:
        tr.setdefault("v", np.zeros_like(tr["mu"]))


def update_v(trials, params, config, session=None):
    if config["method"] != "VB":
        return
    for tr in trials:
        tr.setdefault("w", np.zeros_like(tr["mu"]))
        tr.setdefault("v", np.zeros_like(tr["mu"]))
    with _with_session(trials, params, session) as s:
        s.require_factors()
        nfail = s.ts.update_v()
        if nfail:
            logger.error("Singular I + G'WG (%d systems)", nfail)
        s.pull(trials, ("v",))


def _mstep_dev(s: Session, params, config):
    nfb = s.ts.mstep(config["Mniter"], config["use_hessian"], config["eps"], config["learning_rate"],
                     config["da_bound"], config["db_bound"])
    if nfb:
        logger.error("M-step: %d Newton systems fell back to the gradient step", nfb)
    s.eng.pull_params(params)


def mstep(trials, params, config):
    """Optimise loading and bias (M-step)."""
    if config["Mniter"] < 1:
        return
    with Session(trials, params, upload_factors=False) as s:
        _mstep_dev(s, params, config)


def _hstep_dev(s: Session, trials, params, config):
    from . import gp

    gp._optimize_dev(s, params, config)


def hstep(trials, params, config):
    """GP hyperparameter update (H-step); also refreshes params['cholesky'] like vlgp/gp.py:97."""
    if not config["Hstep"]:
        return
    with Session(trials, params, upload_factors=False) as s:
        _hstep_dev(s, trials, params, config)


# ----------------------------------------------------------------------------------------------------------------------
# constraints
# ----------------------------------------------------------------------------------------------------------------------
def _constrain_loading_dev(s: Session, params, config):
    kind = config["constrain_loading"]
    if not kind or kind == "none":
        return
    a = np.asarray(params["a"], dtype=float)
    L = a.shape[0]
    if kind == "svd":
        _, _, vt = np.linalg.svd(a, full_matrices=False)
        M = a @ vt.T
        params["a"] = vt
    elif kind == "fro":
        sc = np.linalg.norm(a) + config["eps"]
        assign_inplace(params, "a", a / sc)        # in place like the reference (the "svd" branch rebinds there too)
        M = np.eye(L) * sc
    else:
        sc = np.linalg.norm(a, ord=kind, axis=1, keepdims=True) + config["eps"]
        assign_inplace(params, "a", a / sc)
        M = np.diag(sc[:, 0])
    s.eng.push_params(params, which=("a",))
    s.ts.latent_affine(None, M)
    if s.alias is not None and kind != "svd":
        s.ts.latent_affine(None, M, rows=s.alias.shared)          # a shared bin is rescaled once per segment holding it


def _constrain_latent_dev(s: Session, params, config):
    kind = config["constrain_latent"]
    if not kind or kind == "none":
        return
    tot, sq, cnt = s.ts.latent_moments()
    mean = tot / cnt
    std = np.sqrt(np.maximum(sq / cnt - mean * mean, 0.0))
    shift, M = None, None
    if kind in ("location", "both"):
        shift = mean
        b = np.array(params["b"], dtype=float)
        b[0, :] += mean @ params["a"]
        assign_inplace(params, "b", b)
    if kind in ("scale", "both"):
        M = np.diag(1.0 / std)
        assign_inplace(params, "a", np.asarray(params["a"], dtype=float) * std[:, None])
    s.eng.push_params(params, which=("a", "b"))
    if s.alias is None:
        s.ts.latent_affine(shift, M)
    else:       # the reference shifts every segment, then scales every segment: shared bins get each step twice
        if shift is not None:
            s.ts.latent_affine(shift, None)
            s.ts.latent_affine(shift, None, rows=s.alias.shared)
        if M is not None:
            s.ts.latent_affine(None, M)
            s.ts.latent_affine(None, M, rows=s.alias.shared)


def _rebound_keys(config):
    return ("mu",) if config["constrain_loading"] == "svd" else ()


def constrain_loading(trials, params, config):
    """Normalise the loading matrix and rescale the latents accordingly."""
    kind = config["constrain_loading"]
    if not kind or kind == "none":
        return
    with Session(trials, params, upload_factors=False) as s:
        _constrain_loading_dev(s, params, config)
        s.pull(trials, ("mu",), rebind=_rebound_keys(config))


def constrain_latent(trials, params, config):
    """Centre / scale the posterior means and compensate in b / a."""
    kind = config["constrain_latent"]
    if not kind or kind == "none":
        return
    with Session(trials, params, upload_factors=False) as s:
        _constrain_latent_dev(s, params, config)
        s.pull(trials, ("mu",))


# ----------------------------------------------------------------------------------------------------------------------
# outer loop
# ----------------------------------------------------------------------------------------------------------------------
def _estep_aliased(s: Session, config):
    """E-step over segments that share bins, in the reference's order of dependence (see _Aliasing)."""
    ts, al = s.ts, s.alias
    nfail = 0
    for lv, segs in enumerate(al.levels):
        if lv > 0:
            src, dst = al.fwd[lv - 1]
            ts.copy_rows(src, dst, which=("mu", "v"))
        nfail += ts.estep(config["Eniter"], config["dmu_bound"], config["method"], subset=segs)
    ts.copy_rows(al.back_src, al.back_dst, which=("mu", "v"))
    return nfail


def _em_iteration(s: Session, trials, params, config, last=False):
    """One EM iteration on a device session: constrain + E-step, constrain + M-step, H-step (vlgp/core.py:307-326).
    Returns (e_elapsed, m_elapsed, h_elapsed) wall-clock seconds; every stage ends with a device synchronisation."""
    ts = s.ts
    t0 = time.perf_counter()
    s.eng.set_precision(_precision_bits(config))
    _constrain_loading_dev(s, params, config)
    if config["Eniter"] >= 1:
        if s.alias is None:
            nfail = ts.estep(config["Eniter"], config["dmu_bound"], config["method"])
        else:
            nfail = _estep_aliased(s, config)
        if nfail:
            logger.error("E-step: %d r x r systems were not positive definite (update skipped)", nfail)
    t1 = time.perf_counter()
    _constrain_latent_dev(s, params, config)
    if last and os.environ.get("VLGP_PREFETCH", "0") not in ("", "0") and hasattr(ts, "prefetch_state"):
        # vem's last iteration: the posterior is final here (the M- and H-step only read it), so its download can start
        # now on a copy stream and run under them.  w and dmu -- keys the reference rebinds to new arrays -- land in
        # page-locked blocks that pull() hands out as those arrays (no staging copy, no page faults of a fresh block);
        # mu and v land in the context's staging area and are scattered into the caller's arrays by pull().
        # Opt-in (VLGP_PREFETCH=1): measured on B200 boxes it saves 0.8 ms per vem() call (34.6 -> 33.8 ms) once the
        # page-locked blocks exist, but page-locking them (60 MB at config 2) costs ~235 ms the first time -- a loss for
        # a single fit(), a gain only for a loop of hundreds of vem() calls.
        ts.prefetch_state(direct=("w", "dmu") + tuple(_rebound_keys(config)))
    if (config["Mniter"] >= 1 and config["Hstep"] and config.get("overlap_mh", True)
            and not os.environ.get("VLGP_NO_OVERLAP")):
        # The M-step (reads mu, v, y; writes a, b, noise) and the H-step (reads mu, w; writes sigma, omega) of one
        # iteration are independent: the device works through the M-step on a second stream while the host drives
        # the L-BFGS-B rounds of the H-step.  m_elapsed is then the host time the M-step cost on top of the H-step.
        ts.mstep_begin(config["Mniter"], config["use_hessian"], config["eps"], config["learning_rate"],
                       config["da_bound"], config["db_bound"])
        tb = time.perf_counter()
        try:
            _hstep_dev(s, trials, params, config)
        finally:
            tc = time.perf_counter()
            nfb = ts.mstep_end()
        if nfb:
            logger.error("M-step: %d Newton systems fell back to the gradient step", nfb)
        s.eng.pull_params(params)
        t3 = time.perf_counter()
        return t1 - t0, (tb - t1) + (t3 - tc), tc - tb
    if config["Mniter"] >= 1:
        _mstep_dev(s, params, config)
    t2 = time.perf_counter()
    if config["Hstep"]:
        _hstep_dev(s, trials, params, config)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


def vem(trials, params, config, session: Session = None):
    """Variational EM on (already cut) segments; fills config['runtime'] with the reference's keys."""
    callbacks = config["callbacks"]
    tol = config["tol"]
    runtime = {"it": 0, "e_elapsed": [], "m_elapsed": [], "h_elapsed": [], "em_elapsed": []}
    own = session is None
    s = session or Session(trials, params)
    try:
        s.require_factors()
        ts = s.ts
        for it in range(config["max_iter"]):
            runtime["it"] += 1
            norm_mu = np.sqrt(ts.norms()[0])
            norm_a = np.linalg.norm(params["a"])
            norm_b = np.linalg.norm(params["b"])

            te, tm, th = _em_iteration(s, trials, params, config, last=it + 1 >= config["max_iter"])

            runtime["e_elapsed"].append(te)
            runtime["m_elapsed"].append(tm)
            runtime["h_elapsed"].append(th)
            runtime["em_elapsed"].append(te + tm + th)
            config["runtime"] = runtime
            _echo("Iteration {:4d}, E-step {:.2f}s, M-step {:.2f}s".format(runtime["it"], te, tm))

            if callbacks:
                s.pull(trials, rebind=_rebound_keys(config))      # callbacks see coherent host dicts
                for cb in callbacks:
                    try:
                        cb(trials, params, config)
                    except RuntimeError:
                        logger.error("Callback {} failed".format(cb))

            norm_dmu = np.sqrt(ts.norms()[1])
            converged = (norm_dmu < tol * norm_mu and np.linalg.norm(params["da"]) < tol * norm_a
                         and np.linalg.norm(params["db"]) < tol * norm_b)
            if converged and it + 1 >= config["min_iter"]:
                break
        s.pull(trials, rebind=_rebound_keys(config))
    finally:
        if own:
            s.close()
