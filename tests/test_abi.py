"""CPU-only checks of the C-ABI boundary: the shared library builds/loads, exports exactly the symbols that
include/vlgp_b200.h declares, the ctypes table covers all of them, and the product path fails loudly (no CPU
fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vlgp_b200.h")).read()
    return sorted(set(re.findall(r"VLGP_API\s+[\w\s\*]+?\b(vlgp_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from vlgp_b200 import _lib, build

    if not os.path.exists(_lib.lib_path()):
        build.build()
    return _lib.lib_path()


def test_header_symbols_are_exported(libpath):
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(libpath)
    for n in names:
        assert hasattr(lib, n), "libvlgp_b200.so does not export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vlgp_\w+)", out)))
    assert exported == names, "exported symbols differ from the header: %s" % (set(exported) ^ set(names))


def test_ctypes_table_matches_header(libpath):
    from vlgp_b200 import _lib

    assert sorted(_lib.EXPORTS) == _declared()
    _lib.load()


def test_no_cpu_fallback(libpath):
    """Without a CUDA device every entry point raises; nothing is computed on the host."""
    from vlgp_b200 import _lib, engine, core

    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.vlgp_create(0, ctypes.byref(ctx))
    if rc == 0:
        lib.vlgp_destroy(ctx)
        pytest.skip("a GPU is present")
    assert rc < 0 and lib.vlgp_last_error(None)
    engine.reset_engine()
    trials = [dict(y=np.zeros((50, 4)), mu=np.zeros((50, 2)), v=np.zeros((50, 2)), w=np.zeros((50, 2)))]
    params = dict(a=np.zeros((2, 4)), b=np.zeros((1, 4)), noise=np.ones(4), sigma=np.ones(2), omega=np.full(2, 1e-2),
                  likelihood=np.array(["poisson"] * 4), zdim=2, ydim=4, xdim=1, rank=50, gp_noise=1e-4, dt=1)
    from vlgp_b200.preprocess import get_config

    with pytest.raises(_lib.VlgpNativeError):
        core.estep(trials, params, get_config())
    with pytest.raises(_lib.VlgpNativeError):
        core.mstep(trials, params, get_config())


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from vlgp_b200 import _lib

    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("VLGP_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.VlgpNativeError):
        _lib.load()


def test_host_surface_matches_reference_defaults():
    from vlgp_b200.preprocess import get_config, get_params
    from oracle import vlgp_oracle as orc

    cfg = get_config(max_iter=7, bogus=1)
    ref = orc.default_config(max_iter=7)
    extra = {k: cfg.pop(k) for k in list(cfg) if k not in ref}
    assert extra == {"dtype": "float64"}          # the one key the reference does not have (single-precision switch)
    assert cfg == ref and "bogus" not in cfg
    p = get_params([dict(y=np.zeros((10, 4)))], 2, omega_bound=cfg["omega_bound"], lik="gaussian", history=0)
    assert p["xdim"] == 1 and p["rank"] == 50 and list(p["likelihood"]) == ["gaussian"] * 4
    assert np.allclose(p["omega"], 5e-2)


def test_cut_trials_views_and_rng_stream():
    from vlgp_b200.util import cut_trials
    from oracle import vlgp_oracle as orc

    T = 230
    tr = dict(y=np.zeros((T, 3)), x=np.ones((T, 1, 3)), mu=np.arange(T * 2.0).reshape(T, 2), w=np.zeros((T, 2)),
              v=np.zeros((T, 2)))
    np.random.seed(3)
    segs = cut_trials([tr], {}, {"window": 50})
    np.random.seed(3)
    starts = orc.cut_trial_starts(T, 50)
    assert [int(s["mu"][0, 0]) // 2 for s in segs] == starts.tolist()
    assert all(np.shares_memory(s["mu"], tr["mu"]) for s in segs)
    assert segs.dtype == object and all(s["y"].shape == (50, 3) for s in segs)


def test_save_load_roundtrip_and_cli_parser(tmp_path):
    """util.save / util.load keep the reference's file naming (suffix forced to .npy) and drop the unpicklable
    sklearn method; the CLI takes the reference's positional arguments (vlgp/__main__.py:6-11)."""
    from vlgp_b200 import util
    from vlgp_b200.__main__ import cli

    res = {"trials": [{"y": np.zeros((3, 2)), "mu": np.ones((3, 1))}], "params": {"a": np.ones((1, 2)), "transform": print},
           "config": {"window": 50}}
    util.save(res, tmp_path / "out")
    back = util.load(tmp_path / "out.npy")
    assert sorted(back) == ["config", "params", "trials"] and "transform" not in back["params"]
    assert np.array_equal(back["trials"][0]["mu"], np.ones((3, 1)))
    with pytest.raises(FileNotFoundError):
        util.load(tmp_path / "missing.npy")
    with pytest.raises(SystemExit):
        cli(["--help"])


def test_pointer_tables_and_fastpack():
    """Host packing helper: dtype / shape / contiguity are verified in C; mismatches fall back to a converting copy
    (read path) or raise (write path)."""
    from vlgp_b200.engine import _pointer_table

    blocks = [np.zeros((5, 3)), np.ones((2, 3))]
    keep, ptrs, rows = _pointer_table(blocks, np.float64, 3)
    assert np.frombuffer(rows, np.int64).tolist() == [5, 2]
    assert np.frombuffer(ptrs, np.uint64).tolist() == [b.ctypes.data for b in blocks] and keep[0] is blocks[0]
    keep, _, _ = _pointer_table([np.arange(6).reshape(2, 3)], np.float64, 3)          # int64 -> converted copy
    assert keep[0].dtype == np.float64 and keep[0][1, 2] == 5.0
    keep, _, _ = _pointer_table([np.zeros((4, 6))[:, ::2]], np.float64, 3)             # non-contiguous -> copy
    assert keep[0].flags.c_contiguous
    with pytest.raises(ValueError):
        _pointer_table([np.zeros((4, 6))[:, ::2]], np.float64, 3, writable=True)
    with pytest.raises(ValueError):
        _pointer_table([np.zeros((4, 2))], np.float64, 3)


def test_host_count_conversion_is_exact_or_refused(libpath):
    """vlgp_host_f64_to_u8 (the float64 -> uint8 spike-count conversion of the upload path, context-free): exact for
    integer counts in [0, 255] at every length around the vector width, refused (return 0) for anything else at any
    position -- fractions, negatives, 256, NaN, inf, huge values."""
    import ctypes as C

    import numpy as np

    from vlgp_b200 import _lib

    lib = _lib.load()
    assert lib.vlgp_host_pack_isa() in (0, 1)
    u8p = C.POINTER(C.c_ubyte)
    rng = np.random.default_rng(0)
    for n in (1, 3, 15, 16, 17, 31, 32, 33, 1000, 100003):
        y = rng.poisson(0.4, n).astype(float)
        y[rng.integers(n)] = 255.0
        y[rng.integers(n)] = -0.0
        out = np.full(n, 7, np.uint8)
        assert lib.vlgp_host_f64_to_u8(_lib.dptr(y), out.ctypes.data_as(u8p), n) == 1
        assert np.array_equal(out, y.astype(np.uint8)), n
        for bad in (0.5, -1.0, 256.0, 254.99999999, np.nan, np.inf, -np.inf, 1e300, -1e-300, 2.0 ** 31, 2.0 ** 32 + 3):
            for pos in {0, n - 1, n // 2}:
                z = y.copy()
                z[pos] = bad
                assert lib.vlgp_host_f64_to_u8(_lib.dptr(z), out.ctypes.data_as(u8p), n) == 0, (n, bad, pos)
    # unaligned source / destination
    y = rng.poisson(1.0, 1001).astype(float)
    buf = np.zeros(1100, np.uint8)
    assert lib.vlgp_host_f64_to_u8(_lib.dptr(y[1:]), buf[3:].ctypes.data_as(u8p), 1000) == 1
    assert np.array_equal(buf[3:1003], y[1:].astype(np.uint8)) and not buf[:3].any() and not buf[1003:].any()


def test_host_thread_pool_runs_every_task_exactly_once(libpath):
    """The persistent host pool behind the upload / download pipeline (hostpack.cpp): thousands of parallel calls with
    0 .. 64 tasks, from one thread and from several at once, and in a forked child (threads do not survive a fork: the
    pool must notice and rebuild itself instead of waiting for workers that no longer exist)."""
    import ctypes as C
    import os
    import threading

    from vlgp_b200 import _lib

    lib = _lib.load()
    w = C.c_int()
    for n in (0, 1, 2, 3, 7, 8, 16, 17, 64):
        assert lib.vlgp_host_pool_selftest(n, 1500, C.byref(w)) == 0, n
        assert w.value <= 15
    assert lib.vlgp_host_pool_selftest(-1, 1, None) == -1
    out = []
    threads = [threading.Thread(target=lambda k=k: out.append(lib.vlgp_host_pool_selftest(k, 2000, None)))
               for k in (5, 16, 9, 12)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert out == [0, 0, 0, 0]
    pid = os.fork()
    if pid == 0:                                   # child: must not hang on the parent's vanished workers
        rc = 1
        try:
            rc = 0 if lib.vlgp_host_pool_selftest(8, 300, None) == 0 else 1
        finally:
            os._exit(rc)
    import time

    t0 = time.time()
    while True:
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            break
        if time.time() - t0 > 60:
            os.kill(pid, 9)
            os.waitpid(pid, 0)
            raise AssertionError("the forked child hung in the host pool")
        time.sleep(0.05)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
    assert lib.vlgp_host_pool_selftest(8, 300, None) == 0
