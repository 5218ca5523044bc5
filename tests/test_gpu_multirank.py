"""Two-rank SPMD fit() on real GPUs: trials sharded over ranks, allreduce of the M-/H-step statistics inside vem, gather
of the posterior at the end, parity with the reference's golden runs (with and without the H-step).

  * test_spmd_fit_two_ranks_sharing_one_gpu runs on ANY GPU box: both ranks use cuda:0 and exchange through peer-memory
    mailboxes (CUDA IPC) and shared memory only (VLGP_COMM_NO_NCCL=1; NCCL refuses two ranks on one device).  This is
    the in-kernel exchange path of csrc/p2p.cuh -- the one used between GPUs over NVLink -- end to end.
  * test_spmd_fit_two_ranks needs two GPUs: NCCL communicator + peer memory over NVLink; and once more with
    VLGP_NO_P2P=1, NCCL only.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


def _launch(port, **env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "fit_spmd_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("same_on_all_ranks True") == 4, r.stdout[-3000:]
    return r.stdout


@pytest.mark.skipif(_n_gpus() < 1, reason="needs a GPU")
def test_spmd_fit_two_ranks_sharing_one_gpu():
    out = _launch(29531, VLGP_COMM_NO_NCCL="1", VLGP_DEVICE="0")
    assert out.count("peer_memory True") == 4


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_spmd_fit_two_ranks():
    out = _launch(29533)
    assert out.count("peer_memory True") == 4
    out = _launch(29535, VLGP_NO_P2P="1")
    assert out.count("peer_memory False") == 4
