"""CPU check of how engine.TrialSet hands index lists to the C ABI (the row operations behind overlapping windows).

The native entry points need a GPU; the marshalling in front of them does not.  The library is replaced by a recorder
whose functions have the SAME ctypes prototypes as the real ones (vlgp_b200/_lib.py::EXPORTS): every call goes through
ctypes' own argument conversion, and the recorder reads the lists back from the raw pointers it receives."""
import ctypes as C

import numpy as np
import pytest

from vlgp_b200 import _lib, engine as E


class _Recorder:
    """Callable stand-ins built with CFUNCTYPE from the real prototypes; unknown names succeed silently."""

    def __init__(self):
        self.calls = []
        self._keep = []
        for name in ("vlgp_estep_subset", "vlgp_trials_copy_rows", "vlgp_latent_affine_rows", "vlgp_estep",
                     "vlgp_latent_affine"):
            res, args = _lib.EXPORTS[name]
            proto = C.CFUNCTYPE(res, *args)
            fn = proto(getattr(self, "_" + name))
            fn.argtypes = args            # what _lib.load() sets on the real functions
            self._keep.append(fn)
            setattr(self, name, fn)

    def __getattr__(self, name):
        def ok(*a, **k):
            return 0
        return ok

    @staticmethod
    def _arr(ptr, n, dtype):
        return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype).copy() if n else np.zeros(0, dtype)

    def _vlgp_estep(self, ctx, sid, n_iter, bound, vb, nf):
        self.calls.append(("estep", sid, n_iter, bound, vb))
        nf[0] = 0
        return 0

    def _vlgp_estep_subset(self, ctx, sid, n_iter, bound, vb, segs, n, nf):
        self.calls.append(("estep_subset", sid, n_iter, bound, vb, self._arr(segs, n, np.int32)))
        nf[0] = 3
        return 0

    def _vlgp_trials_copy_rows(self, ctx, sid, mask, src, dst, n):
        self.calls.append(("copy_rows", sid, mask, self._arr(src, n, np.int64), self._arr(dst, n, np.int64)))
        return 0

    def _vlgp_latent_affine(self, ctx, sid, shift, M):
        self.calls.append(("affine", sid, bool(shift), bool(M)))
        return 0

    def _vlgp_latent_affine_rows(self, ctx, sid, shift, M, rows, n):
        L = 3
        sh = None if not shift else np.ctypeslib.as_array(shift, shape=(L,)).copy()
        m = None if not M else np.ctypeslib.as_array(M, shape=(L * L,)).copy().reshape(L, L)
        self.calls.append(("affine_rows", sid, sh, m, self._arr(rows, n, np.int64)))
        return 0


@pytest.fixture()
def ts():
    eng = object.__new__(E.Engine)
    eng.lib = _Recorder()
    eng.ctx = C.c_void_p(1)
    eng.device, eng.model_key, eng.N, eng.L, eng.rank = 0, None, 4, 3, 50
    eng.world_size, eng.rank_id, eng.host_allreduce = 1, 0, False
    t = E.TrialSet(eng, [50, 50, 50, 50])
    t.id = 7
    return t


def test_estep_subset_passes_the_list_as_int32(ts):
    rec = ts.eng.lib
    assert ts.estep(4, 2.5, "VB") == 0
    assert rec.calls[-1] == ("estep", 7, 4, 2.5, 1)
    for subset in ([3, 0, 2], np.array([1, 2], dtype=np.int64), np.array([[0], [3]], dtype=np.uint8), range(2)):
        assert ts.estep(25, 5.0, "MAP", subset=subset) == 3           # n_failed comes back through the pointer
        name, sid, n_iter, bound, vb, segs = rec.calls[-1]
        assert (name, sid, n_iter, bound, vb) == ("estep_subset", 7, 25, 5.0, 0)
        assert segs.dtype == np.int32 and np.array_equal(segs, np.asarray(list(np.ravel(subset))))
    ts.estep(1, 5.0, "VB", subset=[])
    assert rec.calls[-1][0] == "estep_subset" and rec.calls[-1][-1].size == 0


def test_copy_rows_mask_and_int64_lists(ts):
    rec = ts.eng.lib
    src, dst = np.array([45, 46, 149], dtype=np.int32), [50, 51, 150]
    ts.copy_rows(src, dst)                                           # default: mu and v
    name, sid, mask, s, d = rec.calls[-1]
    assert (name, sid, mask) == ("copy_rows", 7, 0b0011)
    assert s.dtype == np.int64 and np.array_equal(s, [45, 46, 149]) and np.array_equal(d, [50, 51, 150])
    ts.copy_rows(src[::-1], dst, which=("w", "dmu"))                 # a non-contiguous source list
    assert rec.calls[-1][2] == 0b1100 and np.array_equal(rec.calls[-1][3], [149, 46, 45])
    ts.copy_rows([1], [2], which=("mu", "v", "w", "dmu"))
    assert rec.calls[-1][2] == 0b1111
    with pytest.raises(ValueError):
        ts.copy_rows([1, 2], [3])
    with pytest.raises(KeyError):
        ts.copy_rows([1], [2], which=("y",))


def test_latent_affine_rows_arguments(ts):
    rec = ts.eng.lib
    M = np.arange(9.0).reshape(3, 3)
    ts.latent_affine(None, M)
    assert rec.calls[-1] == ("affine", 7, False, True)
    ts.latent_affine(np.array([1.0, 2.0, 3.0]), None, rows=np.array([5, 199, 0], dtype=np.int16))
    name, sid, sh, m, rows = rec.calls[-1]
    assert name == "affine_rows" and m is None and np.array_equal(sh, [1.0, 2.0, 3.0])
    assert rows.dtype == np.int64 and np.array_equal(rows, [5, 199, 0])
    ts.latent_affine(None, M.T, rows=[7])                            # a transposed (non-contiguous) matrix: row-major copy
    name, sid, sh, m, rows = rec.calls[-1]
    assert sh is None and np.array_equal(m, M.T) and np.array_equal(rows, [7])
    with pytest.raises(ValueError):
        ts.latent_affine(None, np.zeros((2, 2)), rows=[1])


def test_row_ops_switch(ts, monkeypatch):
    monkeypatch.delenv("VLGP_ALIASED_WINDOWS", raising=False)
    assert ts.row_ops is True                                        # the reference's aliased semantics are the default
    monkeypatch.setenv("VLGP_ALIASED_WINDOWS", "0")
    assert ts.row_ops is False
    monkeypatch.setenv("VLGP_ALIASED_WINDOWS", "1")
    assert ts.row_ops is True


def test_pinned_pool_recycles_blocks_with_their_last_view():
    """engine.PinnedPool with malloc / free standing in for the page-locked allocator: an array handed out over a block
    keeps the block out of the pool while ANY view of it lives, the block comes back afterwards and is reused, and a
    closed pool frees instead of keeping."""
    import gc

    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]

    class _Lib:
        allocs, frees = [], []

        def vlgp_host_alloc(self, pp, nbytes):
            addr = libc.malloc(nbytes)
            C.cast(pp, C.POINTER(C.c_void_p))[0] = addr
            self.allocs.append(addr)
            return 0

        def vlgp_host_free(self, p):
            self.frees.append(p.value)
            libc.free(p)
            return 0

    lib = _Lib()
    pool = E.PinnedPool(lib)
    nbytes = 40 * 3 * 8
    addr = pool.take(nbytes)
    assert addr == lib.allocs[0]
    a = pool.as_array(addr, nbytes, (40, 3))
    a[...] = np.arange(120.0).reshape(40, 3)
    views = list(a.reshape(4, 10, 3))
    del a
    gc.collect()
    assert pool.free == {} or not pool.free.get(nbytes)          # views alive: the block is still out
    assert views[3][9, 2] == 119.0
    keep = views[1]
    del views
    gc.collect()
    assert not pool.free.get(nbytes)
    del keep
    gc.collect()
    assert pool.free[nbytes] == [addr]                           # back in the pool
    assert pool.take(nbytes) == addr and len(lib.allocs) == 1    # reused, nothing new allocated
    pool.give_back(addr, nbytes)
    b = pool.as_array(pool.take(nbytes), nbytes, (40, 3))
    pool.close()
    assert lib.frees == []                                       # nothing idle to free ...
    del b
    gc.collect()
    assert lib.frees == [addr]                                   # ... and a block returning to a closed pool is freed
