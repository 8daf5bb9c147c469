"""On-GPU check of the NCCL paths: the row-sharded 8K-style solve on two B200s must equal the single-GPU
image bit for bit, for BOTH schemes (segment scheme of the tridiagonal engine: two small all-reduces; transpose scheme of
the FFT engine: two all-to-alls), also when the Context runs on its own stream and the caller on torch's default stream.
Spawns torchrun (one rank per GPU) over tools/sharded_check.py; skipped on a box with fewer than two GPUs."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("engine,plain", [("tri", False), ("fft", False), ("tri", True), ("fft", True)])
def test_sharded_solve_equals_single_gpu_over_nccl(cuda_lib, engine, plain):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "sharded_check.py"), "--size", "1100", "--engine", engine] + (["--plain-context"] if plain else [])
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, "sharded result differs from the single-GPU result (or the run failed)"
    assert r.stdout.count("sharded == single: True") == 2
