"""Host-side helpers kept from the reference's surface: window cutting (vlgp/util.py:457-499) and the in-place clip
(:446-454).  Everything else in vlgp/util.py is post-hoc analysis and out of scope."""
from __future__ import annotations

import math

import numpy as np

__all__ = ["cut_trials", "cut_trial", "clip", "save", "load"]


def clip(a, lbound, ubound=None):
    """In-place clip; a single positive bound means [-bound, bound]."""
    if ubound is None:
        if not lbound > 0:
            raise AssertionError("bound must be positive")
        lbound, ubound = -lbound, lbound
    elif not ubound > lbound:
        raise AssertionError("ubound must exceed lbound")
    np.clip(a, lbound, ubound, out=a)


_BLAS_CONTROLLER = None


def blas_threads(n):
    """Context manager: cap the BLAS/LAPACK thread pools of this process at ``n`` threads (0 / None: leave them alone).

    The host side of this package only ever touches tiny matrices (L x N loading, the 10 x 10 systems inside scipy's
    L-BFGS-B routine, 15-column randomized SVDs).  A BLAS pool with one thread per core turns each of those calls into
    a thread wake-up: measured here (8 cores, OpenBLAS) one H-step round of five ``setulb`` calls costs 21 ms with the
    default pool and 0.07 ms with one thread -- 300x, and far more than the device work it drives.  The library scan
    behind threadpoolctl is done once per process (about 20 ms), entering the context then costs about 0.1 ms."""
    import contextlib

    global _BLAS_CONTROLLER
    if not n:
        return contextlib.nullcontext()
    try:
        if _BLAS_CONTROLLER is None:
            from threadpoolctl import ThreadpoolController

            # the controller only knows the libraries loaded when it is created: pull in the ones the host path calls
            # into (SciPy ships its own OpenBLAS, loaded with scipy.linalg / scipy.optimize) before taking the snapshot
            import scipy.linalg  # noqa: F401
            import scipy.optimize  # noqa: F401

            _BLAS_CONTROLLER = ThreadpoolController()
        # a cap only ever LOWERS the pool: a process started with OPENBLAS_NUM_THREADS=1 stays at one thread
        current = [lib.num_threads for lib in _BLAS_CONTROLLER.lib_controllers if lib.user_api == "blas"]
        if not current or max(current) <= int(n):
            return contextlib.nullcontext()
        return _BLAS_CONTROLLER.limit(limits=int(n), user_api="blas")
    except Exception:  # pragma: no cover - threadpoolctl missing or an unknown BLAS: run unthrottled
        return contextlib.nullcontext()


def assign_inplace(d, key, value):
    """``d[key] <- value`` keeping the IDENTITY of the array already stored there whenever it can hold the value.

    The reference updates ``params["a"]`` and ``params["b"]`` in place all through ``vem`` (vlgp/core.py:148-149,201,
    219,384-389,406-408), and that is observable: ``initialize`` binds ``params["a"]`` to the very array
    ``FactorAnalysis.components_`` (vlgp/preprocess.py:20,27) and keeps the estimator's bound ``transform`` in
    ``params["transform"]`` (:21), so the map a later ``transform()`` call applies to new trials uses the FITTED loading.
    Rebinding the key instead would silently freeze that map at the FactorAnalysis solution."""
    value = np.asarray(value, dtype=float)
    cur = d.get(key)
    if (isinstance(cur, np.ndarray) and cur.dtype == np.float64 and cur.shape == value.shape and cur.flags.writeable):
        if cur is not value:
            cur[...] = value
    else:
        d[key] = np.array(value, dtype=float)
    return d[key]


def window_starts(length: int, window: int):
    """Start bins of the ceil(length/window) windows covering a trial.  When the length is not a multiple of the
    window, the surplus is spread over the window boundaries by ONE draw of ``np.random.multinomial`` from the global
    RNG (also drawn when the surplus is zero, so the RNG stream matches the reference's)."""
    nseg = math.ceil(length / window)
    surplus = nseg * window - length
    grid = np.arange(nseg, dtype=int) * window
    if nseg > 1:
        spread = np.random.multinomial(surplus, np.full(nseg - 1, 1.0 / (nseg - 1)))
    else:
        # the reference divides by zero here (nseg - 1 == 0) and draws from an empty pvals vector; same draw, no warning
        spread = np.random.multinomial(surplus, np.ones(0))
    return grid - np.concatenate([[0], np.cumsum(spread)])


def cut_trial(trial, window: int):
    """Segments of exactly ``window`` bins whose arrays are VIEWS of the trial's arrays (so in-place updates of a
    segment's mu/v show up in the trial, as in the reference)."""
    segs = []
    for s in window_starts(trial["y"].shape[0], window):
        sl = slice(int(s), int(s) + window)
        segs.append({k: trial[k][sl] for k in ("y", "x", "mu", "w", "v")})
    return segs


def cut_trials(trials, params, config):
    window = config["window"]
    if not window:
        return trials
    out = []
    for tr in trials:
        out.extend(cut_trial(tr, window))
    arr = np.empty(len(out), dtype=object)      # the reference returns an object ndarray (np.concatenate of lists)
    arr[:] = out
    return arr


def save(result, path, ext="npy"):
    """np.save / np.savez of the result dict, same file naming as vlgp/util.py:181-191 (suffix forced to .npy/.npz).
    The bound sklearn method ``params['transform']`` is dropped: it is not picklable across library versions."""
    import pathlib

    path = pathlib.Path(path)
    if isinstance(result, dict) and isinstance(result.get("params"), dict) and "transform" in result["params"]:
        result = dict(result, params={k: v for k, v in result["params"].items() if k != "transform"})
    if ext == "npy":
        np.save(path.with_suffix(".npy"), result, allow_pickle=True)
    elif ext == "npz":
        np.savez(path.with_suffix(".npz"), **result)
    else:
        raise NotImplementedError("unknown file type {}".format(ext))


def load(path):
    """Load a result / trial file written by ``save`` or by the reference (vlgp/util.py:194-208); object arrays need
    ``allow_pickle=True`` on current NumPy, which the reference's loader does not pass."""
    import pathlib

    path = pathlib.Path(path)
    if not path.exists():
        raise FileNotFoundError(path.as_posix())
    if path.suffix == ".npy":
        rez = np.load(path, allow_pickle=True)
        return rez[()] if rez.shape == () else rez
    if path.suffix == ".npz":
        return {**np.load(path, allow_pickle=True)}
    raise NotImplementedError("unknown file type {}".format(path.suffix))
