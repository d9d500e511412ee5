"""N > 1: row-sharded tables with the all-to-all exchange.  The CPU test runs world_size 2 over gloo with
a torch stand-in for the device steps (exchange / routing logic); the GPU test runs the product path
(NCCL + libctr_b200) on 2 GPUs when the box has them."""
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(backend, world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "sharded_worker.py"), backend]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "SHARDED_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_exchange_gloo(world):
    _run("cpu", world)


@pytest.mark.gpu
def test_sharded_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run("cuda", 2)


@pytest.mark.gpu
def test_peer_two_gpus_ipc():
    """The peer-memory path with real CUDA IPC mappings between two processes / two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run("peer", 2)
