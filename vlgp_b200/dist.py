"""Multi-GPU plumbing for the host: one process per GPU (torchrun-style env), trials sharded across ranks, one NCCL
communicator owned by the native engine for the M-/H-step allreduces.

``torch.distributed`` (gloo, CPU) is used ONLY to hand rank 0's 128-byte NCCL unique id to the other ranks and for
host barriers -- plumbing, not the data path.  The data-path collectives are issued by libvlgp_b200.so on its stream.
"""
from __future__ import annotations

import os

import numpy as np

from .engine import get_engine

__all__ = ["init_from_env", "world_size", "rank", "barrier", "shard_bounds", "broadcast_from_root"]

_STATE = {"world": 1, "rank": 0, "pg": False}


def world_size() -> int:
    return _STATE["world"]


def rank() -> int:
    return _STATE["rank"]


def shard_bounds(n_items: int, world: int, r: int):
    """Contiguous, balanced [lo, hi) slice of n_items for rank r of world (first n_items % world ranks get one more)."""
    base, extra = divmod(int(n_items), int(world))
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def init_from_env(backend: str = "gloo"):
    """Initialise from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT.  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    r = int(os.environ.get("RANK", "0"))
    _STATE.update(world=world, rank=r)
    if world == 1:
        return get_engine()
    import torch.distributed as td

    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend=backend, rank=r, world_size=world)
    _STATE["pg"] = True
    eng = get_engine()              # cuda:$LOCAL_RANK
    no_nccl = bool(os.environ.get("VLGP_COMM_NO_NCCL"))   # peer + shared memory only (ranks may then share a GPU)
    box = [(os.urandom(128) if no_nccl else eng.comm_unique_id()) if r == 0 else None]
    td.broadcast_object_list(box, src=0)
    eng.comm_init(r, world, box[0])
    # One node, one process per GPU: scalars that end up on the host are reduced through shared memory (shmcomm.cu).
    # The name is derived from the job's NCCL id, so concurrent jobs on a box do not collide.
    if not os.environ.get("VLGP_NO_SHM") and _single_node(td, world):
        import hashlib

        name = "/vlgp_" + hashlib.sha1(box[0]).hexdigest()[:20]
        eng.attach_host_allreduce(name)
        if not os.environ.get("VLGP_NO_P2P"):
            eng.enable_peer_memory()     # in-kernel allreduce of the M-/H-step statistics over NVLink peer memory
        td.barrier()                     # every rank has mapped the segment: the name can go
        if r == 0:
            try:
                os.unlink("/dev/shm" + name)
            except OSError:
                pass
    if no_nccl and not eng.peer_memory:
        raise RuntimeError("VLGP_COMM_NO_NCCL=1 needs the peer-memory path (one node, shared memory, CUDA IPC)")
    return eng


def _single_node(td, world):
    """True when every rank of the job runs on this host (the shared-memory path needs that)."""
    import socket

    names = [None] * world
    td.all_gather_object(names, socket.gethostname())
    return len(set(names)) == 1


def barrier():
    if _STATE["pg"]:
        import torch.distributed as td

        td.barrier()


def broadcast_from_root(arr):
    """Make every rank hold rank 0's copy of a small host array (sum-allreduce of zeros elsewhere, over NCCL)."""
    eng = get_engine()
    a = np.array(arr, dtype=np.float64, copy=True)
    if eng.world_size > 1:
        if eng.rank_id != 0:
            a[...] = 0.0
        a = eng.allreduce(a)
    return a
