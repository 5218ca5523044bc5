"""Host-side shared-memory allreduce (vlgp_b200/csrc/shmcomm.cu) between several processes of this machine: no GPU is
involved, so the whole protocol -- segment creation, the two-bank sequence numbers, rank-ordered bit-identical sums,
max, ragged lengths, thousands of back-to-back calls -- is exercised by the CPU suite."""
import ctypes as C
import multiprocessing as mp
import os

import numpy as np
import pytest


def _vec(rank, rnd, n):
    rng = np.random.default_rng(1000003 * rnd + rank)
    return rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8)


def _worker(rank, world, name, rounds, q):
    try:
        from vlgp_b200 import _lib

        lib = _lib.load()
        h = C.c_void_p()
        rc = lib.vlgp_shm_open(name.encode(), rank, world, C.byref(h))
        assert rc == 0, "open rc %d" % rc
        dp = C.POINTER(C.c_double)
        for rnd in range(rounds):
            n = 1 + (rnd * 7) % 40 if rnd % 50 else 256
            op = 1 if rnd % 5 == 3 else 0
            buf = _vec(rank, rnd, n).copy()
            rc = lib.vlgp_shm_allreduce(h, buf.ctypes.data_as(dp), n, op)
            assert rc == 0, "allreduce rc %d at round %d" % (rc, rnd)
            if rnd % 97 == 0 or rnd < 8:      # expected value in the library's order: rank 0, 1, 2, ...
                want = _vec(0, rnd, n).copy()
                for r in range(1, world):
                    x = _vec(r, rnd, n)
                    want = np.maximum(want, x) if op else want + x
                assert np.array_equal(buf, want), "round %d mismatch" % rnd
        assert lib.vlgp_shm_allreduce(h, buf.ctypes.data_as(dp), 257, 0) != 0       # too long: rejected, no hang
        lib.vlgp_shm_close(h, 1 if rank == 0 else 0)
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))


@pytest.mark.parametrize("world,rounds", [(2, 3000), (5, 3000), (8, 1500)])
def test_shm_allreduce_between_processes(world, rounds):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "/vlgp_test_%d_%d" % (os.getpid(), world)
    procs = [ctx.Process(target=_worker, args=(r, world, name, rounds, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, msg = q.get(timeout=240)
            results[rank] = msg
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert results == {r: "ok" for r in range(world)}, results
    assert not os.path.exists("/dev/shm" + name)


def test_shm_open_rejects_bad_arguments():
    from vlgp_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    assert lib.vlgp_shm_open(b"no_leading_slash", 0, 2, C.byref(h)) != 0
    assert lib.vlgp_shm_open(b"/vlgp_x", 3, 2, C.byref(h)) != 0
    assert lib.vlgp_shm_open(b"/vlgp_x", 0, 1000, C.byref(h)) != 0
    name = b"/vlgp_test_single_%d" % os.getpid()
    assert lib.vlgp_shm_open(name, 0, 1, C.byref(h)) == 0                      # one rank: allreduce is the identity
    x = np.arange(4.0)
    assert lib.vlgp_shm_allreduce(h, x.ctypes.data_as(C.POINTER(C.c_double)), 4, 0) == 0
    assert np.array_equal(x, np.arange(4.0))
    assert lib.vlgp_shm_close(h, 1) == 0
