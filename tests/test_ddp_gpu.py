"""Multi-rank GPU parity (SURVEY §4 "distributed" tier): needs >= 2 visible GPUs (`gpurun --gpus 2`); the worker is
tests/ddp_gpu_worker.py under torchrun.  Skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_worker(nproc=2, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "ddp_gpu_worker.py")]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_match_mean_of_shard_oracle_gradients():
    res = run_worker(2)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0 and "DDP_GPU_OK" in res.stdout
