"""Host-side handle on the native engine: one ``Engine`` per process/GPU, ``TrialSet`` = device copy of a list of trials.

This is plumbing between the reference's dict surface (trial dicts / params / config, vlgp/api.py:18-76) and the flat
buffers of the C ABI; it does no arithmetic of the hot path itself.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import VlgpNativeError, as_f64, dptr
from .util import assign_inplace

try:  # optional C helper for the pointer tables (host plumbing only); pure-Python fallback below
    if os.environ.get("VLGP_NO_FASTPACK"):      # bench.py --impl reference: none of this package's natives in that process
        raise ImportError("disabled")
    from . import _fastpack
except ImportError:  # pragma: no cover
    _fastpack = None

__all__ = ["Engine", "TrialSet", "get_engine", "reset_engine", "pack_y"]

_ENGINE = None


def get_engine() -> "Engine":
    """Process-wide engine on ``cuda:$LOCAL_RANK`` (created on first use)."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(int(os.environ.get("VLGP_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _ENGINE


def reset_engine():
    global _ENGINE
    if _ENGINE is not None:
        _ENGINE.close()
    _ENGINE = None


def pack_y(ys):
    """Concatenate per-trial observations; store as uint8 when every entry is an integer count in [0, 255] (spike
    counts almost always are), else float64.  Returns (array, ydtype code)."""
    y = ys[0] if len(ys) == 1 else np.concatenate(ys, axis=0)
    if y.dtype == np.uint8:
        return np.ascontiguousarray(y), 1
    yf = np.ascontiguousarray(y, dtype=np.float64)
    if yf.size and yf.min() >= 0 and yf.max() <= 255 and np.array_equal(yf, np.rint(yf)):
        return yf.astype(np.uint8), 1
    return yf, 0


def _pointer_table(arrs, dtype, ncols, writable=False):
    """(arrays kept alive, packed uint64 pointers, packed int64 row counts) of a sequence of 2-D blocks.  Blocks that
    are not C-contiguous arrays of ``dtype`` are converted (copied) first -- except when ``writable``."""
    dt = np.dtype(dtype)
    keep = arrs if isinstance(arrs, list) else list(arrs)
    if _fastpack is not None:
        try:      # dtype / shape / contiguity / writability are all verified in C
            ptrs, rows = _fastpack.pointers(keep, dt.itemsize, ncols, writable, dt.char)
            return keep, ptrs, rows
        except (TypeError, ValueError, BufferError):
            pass
    fixed = []
    for a in keep:
        a = np.asarray(a)
        if a.dtype != dt or not a.flags.c_contiguous or (writable and not a.flags.writeable):
            if writable:
                raise ValueError("expected writable C-contiguous %s blocks" % dt.name)
            a = np.ascontiguousarray(a, dtype=dt)
        if a.ndim != 2 or a.shape[1] != ncols:
            raise ValueError("expected blocks of shape (rows, %d), got %s" % (ncols, a.shape))
        fixed.append(a)
    if _fastpack is not None:
        ptrs, rows = _fastpack.pointers(fixed, dt.itemsize, ncols, writable, dt.char)
    else:
        ptrs = np.array([a.ctypes.data for a in fixed], dtype=np.uint64).tobytes()
        rows = np.array([a.shape[0] for a in fixed], dtype=np.int64).tobytes()
    return fixed, ptrs, rows


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        ctx = C.c_void_p()
        rc = self.lib.vlgp_create(int(device), C.byref(ctx))
        if rc != 0:
            msg = self.lib.vlgp_last_error(None)
            raise VlgpNativeError("vlgp_create(device=%d) failed (%d): %s" % (device, rc, (msg or b"").decode()))
        self.ctx = ctx
        self.device = int(device)
        self.model_key = None
        self.N = self.L = self.rank = 0
        self.xdim = 1
        self.world_size = 1
        self.rank_id = 0
        self.peer_memory = False
        self.pinned = PinnedPool(self.lib)

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def _ck(self, rc, what):
        if rc != 0:
            msg = self.lib.vlgp_last_error(self.ctx)
            raise VlgpNativeError("%s failed (%d): %s" % (what, rc, (msg or b"").decode()))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.vlgp_destroy(self.ctx)
            self.ctx = None
            self.pinned.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        mem = C.c_uint64()
        name = C.create_string_buffer(128)
        self._ck(self.lib.vlgp_device_info(self.ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem), name),
                 "device_info")
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (ma.value, mi.value), "mem": mem.value}

    def sync(self):
        self._ck(self.lib.vlgp_sync(self.ctx), "sync")

    def timer_start(self):
        self._ck(self.lib.vlgp_timer_start(self.ctx), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.vlgp_timer_stop(self.ctx, C.byref(ms)), "timer_stop")
        return float(ms.value)

    def counters(self):
        c = (C.c_int64 * 4)()
        self._ck(self.lib.vlgp_counters(self.ctx, c), "counters")
        return {"launches": c[0], "estep_solves": c[1], "hstep_factorisations": c[2], "collectives": c[3]}

    # -- model -------------------------------------------------------------------------------------------------------
    def ensure_model(self, params):
        """Bind the model dimensions / likelihoods of ``params`` (re-created only when they change)."""
        lik = np.asarray(params["likelihood"])
        mask = np.ascontiguousarray((lik == "poisson").astype(np.uint8))
        bad = ~np.isin(lik, ("poisson", "gaussian"))
        if bad.any():
            raise ValueError("unsupported likelihood(s): %s" % sorted(set(lik[bad].tolist())))
        xdim = int(params.get("xdim", 1) or 1)
        if not 1 <= xdim <= 8:
            raise ValueError("xdim (history) must be between 1 and 8, got %d" % xdim)
        key = (int(params["ydim"]), int(params["zdim"]), int(params["rank"]), mask.tobytes(),
               float(params["gp_noise"]), float(params["dt"]), xdim)
        if key != self.model_key:
            # (vlgp_set_model drops every trial set of the previous model; their handles are refused from then on)
            self._ck(self.lib.vlgp_set_model(self.ctx, key[0], key[1], key[2], mask.ctypes.data_as(_lib.c_u8_p),
                                             key[4], key[5]), "set_model")
            self._ck(self.lib.vlgp_set_regressors(self.ctx, xdim), "set_regressors")
            self.model_key = key
            self.N, self.L, self.rank, self.xdim = key[0], key[1], key[2], xdim

    def push_params(self, params, which=("a", "b", "noise", "sigma", "omega")):
        L, N = self.L, self.N
        a = as_f64(params["a"], (L, N)) if "a" in which else None
        b = as_f64(np.asarray(params["b"]).reshape(self.xdim, -1), (self.xdim, N)) if "b" in which else None
        noise = as_f64(params["noise"], (N,)) if "noise" in which else None
        sigma = as_f64(params["sigma"], (L,)) if "sigma" in which else None
        omega = as_f64(params["omega"], (L,)) if "omega" in which else None
        self._ck(self.lib.vlgp_set_params(self.ctx, dptr(a), dptr(b), dptr(noise), dptr(sigma), dptr(omega)),
                 "set_params")

    def pull_params(self, params, which=("a", "b", "noise", "da", "db")):
        """Copy device parameters into the reference-shaped arrays of ``params`` (b/db are (1, N))."""
        L, N = self.L, self.N
        out = {k: np.empty((L, N)) for k in ("a", "da") if k in which}
        out.update({k: np.empty((self.xdim, N)) for k in ("b", "db") if k in which})
        out.update({k: np.empty((N,)) for k in ("noise",) if k in which})
        g = out.get
        self._ck(self.lib.vlgp_get_params(self.ctx, dptr(g("a")), dptr(g("b")), dptr(g("noise")), dptr(g("da")),
                                          dptr(g("db")), None, None), "get_params")
        for k, val in out.items():
            if k == "noise":
                params[k] = val              # the reference rebinds noise (vlgp/core.py:177,244) ...
            else:
                assign_inplace(params, k, val)   # ... and updates a, b, da, db in place (:148-149,155-156,201,219)
        return params

    def new_trials(self, lengths) -> "TrialSet":
        return TrialSet(self, lengths)

    # -- communicator ------------------------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.lib.vlgp_comm_unique_id(self.ctx, _lib.nccl_path().encode(), buf), "comm_unique_id")
        return buf.raw

    def comm_init(self, rank: int, world_size: int, unique_id: bytes):
        if world_size > 1:
            assert len(unique_id) == 128
            self._ck(self.lib.vlgp_comm_init(self.ctx, _lib.nccl_path().encode(), int(rank), int(world_size),
                                             C.create_string_buffer(unique_id, 128)), "comm_init")
        self.world_size, self.rank_id = int(world_size), int(rank)

    def attach_host_allreduce(self, name: str):
        """Open the job's shared-memory segment (same ``name`` on every rank of the node) and hand it to the context:
        host-consumed scalars (H-step objective values, norms) are then summed on the host instead of through NCCL."""
        if self.world_size <= 1:
            return
        h = C.c_void_p()
        rc = self.lib.vlgp_shm_open(name.encode(), self.rank_id, self.world_size, C.byref(h))
        if rc:
            raise _lib.VlgpNativeError("vlgp_shm_open(%s) failed with status %d" % (name, rc))
        self._ck(self.lib.vlgp_comm_attach_shm(self.ctx, h), "comm_attach_shm")
        self.host_allreduce = True

    def enable_peer_memory(self) -> bool:
        """Map every rank's mailbox (collective; after attach_host_allreduce): the small device-side reductions then go
        through peer memory from inside the kernels (csrc/p2p.cuh).  False when some rank cannot map some peer."""
        if self.world_size <= 1:
            return False
        on = C.c_int()
        self._ck(self.lib.vlgp_comm_enable_p2p(self.ctx, C.byref(on)), "comm_enable_p2p")
        self.peer_memory = bool(on.value)
        return self.peer_memory

    def allreduce(self, x, op="sum"):
        """In-place allreduce of a small host array (<= 256 doubles per call; chunked here)."""
        a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if self.world_size > 1:
            for i in range(0, a.size, 256):
                chunk = np.ascontiguousarray(a[i:i + 256])
                self._ck(self.lib.vlgp_comm_allreduce(self.ctx, dptr(chunk), chunk.size, 1 if op == "max" else 0),
                         "comm_allreduce")
                a[i:i + 256] = chunk
        return a.reshape(np.shape(x))

    def allreduce_bulk(self, a):
        """In-place sum-allreduce of a large C-contiguous float64 host array."""
        if self.world_size > 1:
            assert a.dtype == np.float64 and a.flags.c_contiguous
            self._ck(self.lib.vlgp_comm_allreduce_bulk(self.ctx, dptr(a), int(a.size)), "comm_allreduce_bulk")
        return a

    # -- measurement -------------------------------------------------------------------------------------------------
    def peak_fp64(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.vlgp_peak_fp64(self.ctx, C.byref(a), C.byref(b)), "peak_fp64")
        return {"dfma_tflops": a.value, "dmma_tflops": b.value}

    def peak_hbm(self, nbytes=1 << 30):
        g = C.c_double()
        self._ck(self.lib.vlgp_peak_hbm(self.ctx, int(nbytes), C.byref(g)), "peak_hbm")
        return g.value

    def flush_l2(self):
        self._ck(self.lib.vlgp_flush_l2(self.ctx), "flush_l2")

    def set_precision(self, bits=64):
        """Arithmetic of the segment E-step's rate passes: 64 (default, reference-exact) or 32 (include/vlgp_b200.h)."""
        self._ck(self.lib.vlgp_set_precision(self.ctx, int(bits)), "set_precision")
        self.precision = int(bits)

    def profile_enable(self, mask=0xF):
        """Time kernel classes with CUDA events: bit 0 E-step, 1 M-step statistics, 2 H-step segments, 3 ichol."""
        self._ck(self.lib.vlgp_profile_enable(self.ctx, int(mask)), "profile_enable")

    def profile_get(self, which):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.lib.vlgp_profile_get(self.ctx, int(which), C.byref(ms), C.byref(n)), "profile_get")
        return ms.value, n.value


class PinnedPool:
    """Page-locked host blocks that a state prefetch fills directly and that are then handed out AS the arrays of
    trial["w"] / trial["dmu"] (the reference rebinds those keys to new arrays in every E-step, vlgp/core.py:117-120):
    the download of those two arrays costs the host nothing -- no staging copy, no first-touch page faults of a fresh
    np.empty block.  A block returns to the pool when the last NumPy view of it is garbage-collected (a finalizer on the
    ctypes buffer every view has as its base), so a loop of vem() calls cycles through the same two or four blocks."""

    MAX_KEPT = 8

    def __init__(self, lib):
        self.lib = lib
        self.free = {}              # nbytes -> [address, ...]
        self.closed = False

    def take(self, nbytes):
        """Address of a page-locked block of ``nbytes`` bytes, or None when it cannot be had."""
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        p = C.c_void_p()
        if self.lib.vlgp_host_alloc(C.byref(p), int(nbytes)) != 0 or not p.value:
            return None
        return p.value

    def give_back(self, addr, nbytes):
        lst = self.free.setdefault(nbytes, [])
        if self.closed or sum(len(v) for v in self.free.values()) >= self.MAX_KEPT:
            try:
                self.lib.vlgp_host_free(C.c_void_p(addr))
            except Exception:      # noqa: BLE001 -- interpreter shutdown
                pass
        else:
            lst.append(addr)

    def as_array(self, addr, nbytes, shape):
        """float64 array over the block; the block goes back to the pool when the array and all its views are gone."""
        import weakref

        buf = (C.c_char * nbytes).from_address(addr)
        weakref.finalize(buf, self.give_back, addr, nbytes)
        return np.frombuffer(buf, dtype=np.float64).reshape(shape)

    def close(self):
        self.closed = True
        for lst in self.free.values():
            for addr in lst:
                self.lib.vlgp_host_free(C.c_void_p(addr))
        self.free = {}


class TrialSet:
    """Device-resident copy of a list of trials (or segments): y, mu, v, w, dmu and the prior factors per length."""

    def __init__(self, eng: Engine, lengths):
        self.eng = eng
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        if self.lengths.ndim != 1 or self.lengths.size == 0:
            raise ValueError("lengths must be a non-empty 1-D sequence")
        self.nbin = int(self.lengths.sum())
        self.starts = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)
        sid = C.c_int()
        eng._ck(eng.lib.vlgp_trials_create(eng.ctx, int(self.lengths.size), self.lengths.ctypes.data_as(_lib.c_i32_p),
                                           C.byref(sid)), "trials_create")
        self.id = sid.value
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.y_stored = None          # dtype code of y in HBM once uploaded (1 = uint8 counts, 0 = float64)
        self.general_x = False        # True once regressors other than the all-ones bias column were uploaded
        self._pinned = {}             # key -> address of the page-locked block a prefetch is filling / has filled

    def free(self):
        if self.id is not None and self.eng.ctx:
            self.eng._ck(self.eng.lib.vlgp_trials_free(self.eng.ctx, self.id), "trials_free")   # waits for a prefetch
        self.id = None
        self._drop_pinned()

    def _drop_pinned(self):
        nbytes = self.nbin * self.eng.L * 8
        for addr in self._pinned.values():
            self.eng.pinned.give_back(addr, nbytes)
        self._pinned = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()
        return False

    def _lib(self):
        return self.eng.lib, self.eng.ctx

    # -- data ----------------------------------------------------------------------------------------------------------
    def set_y(self, y, ydtype=None):
        lib, ctx = self._lib()
        if ydtype is None:
            y, ydtype = pack_y([y])
        y = np.ascontiguousarray(y)
        if y.shape != (self.nbin, self.eng.N):
            raise ValueError("y must be (%d, %d), got %s" % (self.nbin, self.eng.N, y.shape))
        self.eng._ck(lib.vlgp_trials_set_y(ctx, self.id, y.ctypes.data_as(C.c_void_p), int(ydtype)), "trials_set_y")
        self.h2d_bytes += y.nbytes
        self.y_stored = int(ydtype)

    def set_y_parts(self, ys):
        """Upload per-trial observation blocks without concatenating them on the host (native pinned pipeline).
        Returns the stored dtype code (1 = uint8 counts, 0 = float64)."""
        lib, ctx = self._lib()
        N = self.eng.N
        if len(ys) != self.lengths.size:
            raise ValueError("expected %d observation blocks, got %d" % (self.lengths.size, len(ys)))
        src_u8 = isinstance(ys[0], np.ndarray) and ys[0].dtype == np.uint8     # mixed dtypes fall to the slow path
        want = np.uint8 if src_u8 else np.float64
        keep, ptrs, rows = _pointer_table(ys, want, N)
        if not np.array_equal(np.frombuffer(rows, dtype=np.int64), self.lengths):
            raise ValueError("observation blocks do not match the trial lengths")
        stored = C.c_int()
        self.eng._ck(lib.vlgp_trials_set_y_parts(ctx, self.id, len(keep), ptrs, rows, 1 if src_u8 else 0,
                                                 C.byref(stored)), "trials_set_y_parts")
        self.h2d_bytes += self.nbin * N * (1 if stored.value == 1 else 8)
        self.y_stored = stored.value
        return stored.value

    def set_x(self, x):
        """General regressors: the trials' x blocks concatenated, (nbin, xdim, N) float64 (csrc/regress.cu)."""
        lib, ctx = self._lib()
        x = as_f64(x, (self.nbin, self.eng.xdim, self.eng.N))
        self.eng._ck(lib.vlgp_trials_set_x(ctx, self.id, dptr(x)), "trials_set_x")
        self.h2d_bytes += x.nbytes
        self.general_x = True

    _WHICH = {"mu": 0, "v": 1, "w": 2, "dmu": 3}

    def set_state_parts(self, **blocks):
        """Upload mu / v / w given as one (rows_i, L) float64 block per trial, gathered natively (no host concat)."""
        lib, ctx = self._lib()
        for key, arrs in blocks.items():
            if arrs is None:
                continue
            keep, ptrs, rows = _pointer_table(arrs, np.float64, self.eng.L)
            self.eng._ck(lib.vlgp_trials_set_state_parts(ctx, self.id, self._WHICH[key], len(keep), ptrs, rows),
                         "trials_set_state_parts")
            self.h2d_bytes += self.nbin * self.eng.L * 8

    def get_state_parts(self, **blocks):
        """Download mu / v / w / dmu straight INTO the given per-trial float64 blocks (in place)."""
        lib, ctx = self._lib()
        for key, arrs in blocks.items():
            if arrs is None:
                continue
            keep, ptrs, rows = _pointer_table(arrs, np.float64, self.eng.L, writable=True)
            self.eng._ck(lib.vlgp_trials_get_state_parts(ctx, self.id, self._WHICH[key], len(keep), ptrs, rows),
                         "trials_get_state_parts")
            self.d2h_bytes += self.nbin * self.eng.L * 8

    def project_y(self, mean, P, Cz):
        """mu <- ((y - mean) @ P) @ Cz on the device: FactorAnalysis.transform of every bin (vlgp/preprocess.py:36-41;
        P = Wpsi' (N x L), Cz = cov_z (L x L) of the fitted factor model)."""
        lib, ctx = self._lib()
        mean = as_f64(mean, (self.eng.N,))
        P = as_f64(P, (self.eng.N, self.eng.L))
        Cz = as_f64(Cz, (self.eng.L, self.eng.L))
        self.eng._ck(lib.vlgp_trials_project_y(ctx, self.id, dptr(mean), dptr(P), dptr(Cz)), "trials_project_y")
        self.h2d_bytes += mean.nbytes + P.nbytes + Cz.nbytes

    def set_state(self, mu=None, v=None, w=None):
        lib, ctx = self._lib()
        shp = (self.nbin, self.eng.L)
        arrs = [None if x is None else as_f64(x, shp) for x in (mu, v, w)]
        self.eng._ck(lib.vlgp_trials_set_state(ctx, self.id, *[dptr(x) for x in arrs]), "trials_set_state")
        self.h2d_bytes += sum(x.nbytes for x in arrs if x is not None)

    def get_state(self, which=("mu", "v", "w", "dmu")):
        lib, ctx = self._lib()
        shp = (self.nbin, self.eng.L)
        out = {k: np.empty(shp) for k in which}
        g = out.get
        self.eng._ck(lib.vlgp_trials_get_state(ctx, self.id, dptr(g("mu")), dptr(g("v")), dptr(g("w")), dptr(g("dmu"))),
                     "trials_get_state")
        self.d2h_bytes += sum(x.nbytes for x in out.values())
        return out

    # -- prior factor -------------------------------------------------------------------------------------------------
    def make_cholesky(self):
        lib, ctx = self._lib()
        self.eng._ck(lib.vlgp_make_cholesky(ctx, self.id), "make_cholesky")

    def get_cholesky(self, length, with_pivots=False):
        lib, ctx = self._lib()
        L, r = self.eng.L, self.eng.rank
        G = np.empty((L, int(length), r))
        piv = np.empty((L, r), dtype=np.int32)
        ncol = np.empty((L,), dtype=np.int32)
        self.eng._ck(lib.vlgp_get_cholesky(ctx, self.id, int(length), dptr(G), piv.ctypes.data_as(_lib.c_i32_p),
                                           ncol.ctypes.data_as(_lib.c_i32_p)), "get_cholesky")
        return (G, piv, ncol) if with_pivots else G

    def set_cholesky(self, length, G):
        lib, ctx = self._lib()
        G = as_f64(G, (self.eng.L, int(length), self.eng.rank))
        self.eng._ck(lib.vlgp_set_cholesky(ctx, self.id, int(length), dptr(G)), "set_cholesky")
        self.h2d_bytes += G.nbytes

    # -- steps -----------------------------------------------------------------------------------------------------------
    @property
    def row_ops(self):
        """True when the host drives the reference's in-place semantics of OVERLAPPING windows through
        ``estep(subset=)``, ``copy_rows`` and ``latent_affine(rows=)`` (vlgp_b200/core.py::_Aliasing): the default.
        VLGP_ALIASED_WINDOWS=0 treats overlapping windows as independent copies instead (one batched E-step, last
        writer wins on the shared bins -- NOT what the reference computes)."""
        return os.environ.get("VLGP_ALIASED_WINDOWS", "1") not in ("", "0")

    def estep(self, n_iter, dmu_bound=5.0, method="VB", subset=None):
        """E-step on every member of the set, or on the listed members only (``subset``: indices, each once)."""
        lib, ctx = self._lib()
        nf = C.c_int()
        if subset is None:
            self.eng._ck(lib.vlgp_estep(ctx, self.id, int(n_iter), float(dmu_bound), int(method == "VB"),
                                        C.byref(nf)), "estep")
        else:
            sub = np.ascontiguousarray(subset, dtype=np.int32).reshape(-1)
            self.eng._ck(lib.vlgp_estep_subset(ctx, self.id, int(n_iter), float(dmu_bound), int(method == "VB"),
                                               sub.ctypes.data_as(_lib.c_i32_p), int(sub.size), C.byref(nf)),
                         "estep_subset")
        return nf.value

    def copy_rows(self, src, dst, which=("mu", "v")):
        """bin dst[i] <- bin src[i] (indices into the set's concatenated bins) in the listed per-bin arrays."""
        lib, ctx = self._lib()
        src = np.ascontiguousarray(src, dtype=np.int64).reshape(-1)
        dst = np.ascontiguousarray(dst, dtype=np.int64).reshape(-1)
        if src.size != dst.size:
            raise ValueError("copy_rows: %d sources, %d destinations" % (src.size, dst.size))
        mask = 0
        for k in which:
            mask |= 1 << self._WHICH[k]
        self.eng._ck(lib.vlgp_trials_copy_rows(ctx, self.id, mask, src.ctypes.data_as(_lib.c_i64_p),
                                               dst.ctypes.data_as(_lib.c_i64_p), int(src.size)), "trials_copy_rows")

    def update_w(self):
        lib, ctx = self._lib()
        self.eng._ck(lib.vlgp_update_w(ctx, self.id), "update_w")

    def update_v(self):
        lib, ctx = self._lib()
        nf = C.c_int()
        self.eng._ck(lib.vlgp_update_v(ctx, self.id, C.byref(nf)), "update_v")
        return nf.value

    def mstep(self, n_iter, use_hessian=True, eps=1e-8, learning_rate=1.0, da_bound=5.0, db_bound=5.0):
        lib, ctx = self._lib()
        nf = C.c_int()
        self.eng._ck(lib.vlgp_mstep(ctx, self.id, int(n_iter), int(bool(use_hessian)), float(eps), float(learning_rate),
                                    float(da_bound), float(db_bound), C.byref(nf)), "mstep")
        return nf.value

    def mstep_begin(self, n_iter, use_hessian=True, eps=1e-8, learning_rate=1.0, da_bound=5.0, db_bound=5.0):
        """Enqueue the M-step on its own stream and return; ``mstep_end`` waits for it (H-step runs in between)."""
        lib, ctx = self._lib()
        self.eng._ck(lib.vlgp_mstep_begin(ctx, self.id, int(n_iter), int(bool(use_hessian)), float(eps),
                                          float(learning_rate), float(da_bound), float(db_bound)), "mstep_begin")

    def mstep_end(self):
        lib, ctx = self._lib()
        nf = C.c_int()
        self.eng._ck(lib.vlgp_mstep_end(ctx, C.byref(nf)), "mstep_end")
        return nf.value

    def hstep_prepare(self):
        lib, ctx = self._lib()
        self.eng._ck(lib.vlgp_hstep_prepare(ctx, self.id), "hstep_prepare")

    def hstep_objective(self, latent, hyper):
        """(ll, dll/dlog omega, info) at hyper = (sigma^2, omega, eps)."""
        lib, ctx = self._lib()
        h = as_f64(hyper, (3,))
        ll, dll, info = C.c_double(), C.c_double(), C.c_int()
        self.eng._ck(lib.vlgp_hstep_objective(ctx, self.id, int(latent), dptr(h), C.byref(ll), C.byref(dll),
                                              C.byref(info)), "hstep_objective")
        return ll.value, dll.value, info.value

    def hstep_objective_batch(self, latents, hypers):
        """Evaluate several (latent, hyper) pairs in one device pass; returns (ll[n], dll[n], info[n])."""
        lib, ctx = self._lib()
        lat = np.ascontiguousarray(latents, dtype=np.int32)
        h = as_f64(hypers, (lat.size, 3))
        ll, dll = np.empty(lat.size), np.empty(lat.size)
        info = np.empty(lat.size, dtype=np.int32)
        self.eng._ck(lib.vlgp_hstep_objective_batch(ctx, self.id, int(lat.size), lat.ctypes.data_as(_lib.c_i32_p),
                                                    dptr(h), dptr(ll), dptr(dll), info.ctypes.data_as(_lib.c_i32_p)),
                     "hstep_objective_batch")
        return ll, dll, info

    def hstep_optimize(self, latents, initials, bounds, mask=(0, 1, 0), collapse_tol=1e-9):
        """L-BFGS-B over log(sigma^2, omega, eps) of the listed latents, all rounds inside ONE native call
        (vlgp_hstep_optimize; includes hstep_prepare).  ``initials``: (n, 3) and ``bounds``: (3, 2) in natural units, like
        the reference passes them to optimze1d (vlgp/gp.py:83-90).  Returns (hyper (n, 3) natural units, fval (n,),
        nfev (n,), task (n,), device rounds)."""
        lib, ctx = self._lib()
        lat = np.ascontiguousarray(latents, dtype=np.int32)
        n = lat.size
        x0 = as_f64(np.log(np.asarray(initials, dtype=np.float64)), (n, 3))
        lb = as_f64(np.log(np.asarray(bounds, dtype=np.float64)), (3, 2))
        mk = np.ascontiguousarray(mask, dtype=np.int32)
        if mk.shape != (3,):
            raise ValueError("mask must have 3 entries")
        res, fval = np.empty((n, 3)), np.empty(n)
        nfev, task = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        rounds = C.c_int32()
        self.eng._ck(lib.vlgp_hstep_optimize(ctx, self.id, int(n), lat.ctypes.data_as(_lib.c_i32_p), dptr(x0), dptr(lb),
                                             mk.ctypes.data_as(_lib.c_i32_p), float(collapse_tol), dptr(res), dptr(fval),
                                             nfev.ctypes.data_as(_lib.c_i32_p), task.ctypes.data_as(_lib.c_i32_p),
                                             C.byref(rounds)), "hstep_optimize")
        return np.exp(res), fval, nfev, task, rounds.value

    def gpfa_estep(self, C, d, rho, P):
        """mu <- P h for every (equal-length) segment, h = bigC' bigR^-1 (y - d) (vlgp/gpfa.py:37-45; csrc/gpfa.cu)."""
        lib, ctx = self._lib()
        C = np.ascontiguousarray(C, dtype=np.float64)
        d = np.ascontiguousarray(np.ravel(d), dtype=np.float64)
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        PT = np.ascontiguousarray(np.asarray(P, dtype=np.float64).T)
        self.eng._ck(lib.vlgp_gpfa_estep(ctx, self.id, dptr(C), dptr(d), dptr(rho), dptr(PT)), "gpfa_estep")

    def gpfa_stats(self):
        """(Z1'Z1, Z1'Y, sum y^2) with Z1 = [mu, 1] over all bins of the set (csrc/gpfa.cu)."""
        lib, ctx = self._lib()
        L1, N = self.eng.L + 1, self.eng.N
        ztz, zty, yy = np.empty((L1, L1)), np.empty((L1, N)), np.empty(N)
        self.eng._ck(lib.vlgp_gpfa_stats(ctx, self.id, dptr(ztz), dptr(zty), dptr(yy)), "gpfa_stats")
        return ztz, zty, yy

    def prefetch_state(self, which=("mu", "v", "w", "dmu"), direct=()):
        """Start the device-to-host copy of the listed state arrays behind what is enqueued so far; a following
        get_state_parts is served from it unless the state was written in between (include/vlgp_b200.h).  Arrays listed
        in ``direct`` go straight into page-locked blocks of their own that ``take_prefetched`` turns into the result
        arrays (no further host copy)."""
        lib, ctx = self._lib()
        names = ("mu", "v", "w", "dmu")
        mask = sum(1 << names.index(k) for k in which)
        if self._pinned:                                   # an earlier prefetch of this set: wait for it, recycle
            self.eng._ck(lib.vlgp_trials_prefetch_wait(ctx, self.id), "prefetch_wait")
            self._drop_pinned()
        nbytes = self.nbin * self.eng.L * 8
        dst = (C.c_void_p * 4)()
        for k in direct:
            if k in which:
                addr = self.eng.pinned.take(nbytes)
                if addr is not None:
                    self._pinned[k] = addr
                    dst[names.index(k)] = addr
        self.eng._ck(lib.vlgp_trials_prefetch_state_into(ctx, self.id, int(mask), dst), "prefetch_state")

    def take_prefetched(self, key):
        """(nbin, L) array over the page-locked block a prefetch filled with the CURRENT array ``key``, or None (no
        such prefetch, or the state has been written since).  The caller owns the array; the block returns to the pool
        with its last view."""
        addr = self._pinned.get(key)
        if addr is None or self.id is None:
            return None
        lib, ctx = self._lib()
        valid = C.c_int()
        self.eng._ck(lib.vlgp_trials_prefetch_take(ctx, self.id, ("mu", "v", "w", "dmu").index(key), C.byref(valid)),
                     "prefetch_take")
        if not valid.value:
            return None
        del self._pinned[key]
        nbytes = self.nbin * self.eng.L * 8
        self.d2h_bytes += nbytes
        return self.eng.pinned.as_array(addr, nbytes, (self.nbin, self.eng.L))

    def posterior_cov(self, trial, latent, reg=1e-6):
        """(T, T) posterior covariance inv(inv(G G' + reg I) + diag(w)) of one latent of one member of the set."""
        lib, ctx = self._lib()
        T = int(self.lengths[int(trial)])
        cov = np.empty((T, T))
        self.eng._ck(lib.vlgp_posterior_cov(ctx, self.id, int(trial), int(latent), float(reg), dptr(cov)), "posterior_cov")
        self.d2h_bytes += cov.nbytes
        return cov

    def latent_affine(self, shift=None, M=None, rows=None):
        """mu <- (mu - shift) @ M on every bin, or on the listed bins only (``rows``: bin indices, each once)."""
        lib, ctx = self._lib()
        L = self.eng.L
        s = None if shift is None else as_f64(np.asarray(shift).reshape(-1), (L,))
        m = None if M is None else as_f64(M, (L, L))
        if rows is None:
            self.eng._ck(lib.vlgp_latent_affine(ctx, self.id, dptr(s), dptr(m)), "latent_affine")
        else:
            r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
            self.eng._ck(lib.vlgp_latent_affine_rows(ctx, self.id, dptr(s), dptr(m), r.ctypes.data_as(_lib.c_i64_p),
                                                     int(r.size)), "latent_affine_rows")

    def norms(self):
        """(sum mu^2, sum dmu^2) over all bins (all ranks)."""
        lib, ctx = self._lib()
        out = np.empty(2)
        self.eng._ck(lib.vlgp_norms(ctx, self.id, dptr(out)), "norms")
        return float(out[0]), float(out[1])

    def latent_moments(self):
        lib, ctx = self._lib()
        L = self.eng.L
        s, q = np.empty(L), np.empty(L)
        n = C.c_int64()
        self.eng._ck(lib.vlgp_latent_moments(ctx, self.id, dptr(s), dptr(q), C.byref(n)), "latent_moments")
        return s, q, n.value
