"""World-size-2 CPU (gloo) tests of the N > 1 host logic: trial sharding, and the fact the engine relies on -- the
M-step / H-step sufficient statistics are sums over trials, so per-rank statistics + one sum-allreduce reproduce the
single-process update.  The per-rank arithmetic here is the NumPy oracle (no GPU in this container); on the GPU box the
same allreduce is NCCL inside libvlgp_b200.so (csrc/comm.cu)."""
import os
import socket

import numpy as np
import pytest

from oracle import vlgp_oracle as orc
from vlgp_b200.dist import shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mstep_stats(y, mu, v, a, b):
    """Per-neuron Poisson sufficient statistics of one Newton iteration, as mstep_stats_kernel lays them out:
    grad_a (L), packed lower Hessian (L(L+1)/2), grad_b, hess_b, sum e, sum e^2 (vlgp/core.py:174-189,205-208)."""
    L, N = a.shape
    eta = mu @ a + b
    r = orc.trunc_exp(eta + 0.5 * (v @ (a * a)))
    out = []
    for n in range(N):
        s = mu + v * a[:, n]
        g = mu.T @ y[:, n] - s.T @ r[:, n]
        H = s.T @ (r[:, [n]] * s)
        H[np.diag_indices(L)] += r[:, n] @ v
        e = y[:, n] - eta[:, n]
        out.append(np.concatenate([g, H[np.tril_indices(L)], [np.sum(y[:, n] - r[:, n]), np.sum(r[:, n]),
                                                                 e.sum(), (e * e).sum()]]))
    return np.array(out).T          # nstat x N


def _mstep_solve(stat, count, a, b, eps=1e-8, bound=5.0):
    L, N = a.shape
    a, b = a.copy(), b.copy()
    for n in range(N):
        g = stat[:L, n]
        H = np.zeros((L, L))
        H[np.tril_indices(L)] = stat[L:L + L * (L + 1) // 2, n]
        H = H + np.tril(H, -1).T
        a[:, n] += np.clip(np.linalg.solve(H + eps * np.eye(L), g), -bound, bound)
        b[0, n] += np.clip(stat[-4, n] / (stat[-3, n] + eps), -bound, bound)
    me = stat[-2] / count
    return a, b, stat[-1] / count - me * me


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)               # every rank builds the same global problem, keeps its shard
    n_trials, T, N, L = 6, 50, 7, 2
    y = rng.poisson(0.3, (n_trials, T, N)).astype(float)
    mu = 0.3 * rng.standard_normal((n_trials, T, L))
    v = 0.05 * rng.random((n_trials, T, L))
    a = 0.3 * rng.standard_normal((L, N))
    b = np.full((1, N), -1.0)
    lo, hi = shard_bounds(n_trials, world, rank)
    for _ in range(3):                            # three Newton iterations, one allreduce each
        stat = _mstep_stats(y[lo:hi].reshape(-1, N), mu[lo:hi].reshape(-1, L), v[lo:hi].reshape(-1, L), a, b)
        buf = torch.from_numpy(np.concatenate([stat.ravel(), [float((hi - lo) * T)]]))
        td.all_reduce(buf)
        tot = buf.numpy()
        a, b, noise = _mstep_solve(tot[:-1].reshape(stat.shape), tot[-1], a, b)
    td.barrier()
    td.destroy_process_group()
    q.put((rank, a, b, noise))


def test_sharded_mstep_matches_single_process():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks hold identical parameters
    for k in (1, 2, 3):
        assert np.array_equal(res[0][k], res[1][k])
    # and they equal the oracle's unsharded M-step on the union of the trials
    rng = np.random.default_rng(0)
    n_trials, T, N, L = 6, 50, 7, 2
    y = rng.poisson(0.3, (n_trials, T, N)).astype(float)
    mu = 0.3 * rng.standard_normal((n_trials, T, L))
    v = 0.05 * rng.random((n_trials, T, L))
    a = 0.3 * rng.standard_normal((L, N))
    b = np.full((1, N), -1.0)
    a1, b1, noise1, _, _ = orc.mstep_arrays(y.reshape(-1, N), np.ones((n_trials * T, 1, N)), mu.reshape(-1, L),
                                            v.reshape(-1, L), a, b, np.ones(N, bool), 3)
    assert np.max(np.abs(res[0][1] - a1)) < 1e-11 * np.max(np.abs(a1))
    assert np.max(np.abs(res[0][2] - b1)) < 1e-11 * np.max(np.abs(b1))
    assert np.max(np.abs(res[0][3] - noise1)) < 1e-11 * np.max(np.abs(noise1))
