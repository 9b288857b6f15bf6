"""GPU, needs >= 2 devices (skipped on a one-GPU box): the fused power-map + all-gather kernel
(epilogue stores into every rank's buffer over NVLink peer memory, step flags inside the kernel)
against a one-GPU launch, bit for bit, on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_fused_peer_gather_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "_peer_gather_child.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "PEER_GATHER_OK" in out.stdout, (out.stdout[-1000:], out.stderr[-2000:])


def test_fd_sharded_two_ranks():
    """Frequency-domain maps with the directions sharded over two GPUs (MVDR on the tcgen05 kernel and DAS):
    assembled maps on every rank == the one-GPU maps, bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29543", os.path.join(HERE, "_fd_sharded_child.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "FD_SHARDED_OK" in out.stdout, (out.stdout[-1000:], out.stderr[-2000:])
