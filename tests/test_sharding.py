"""Sharded (multi-GPU) path: host logic under gloo on CPU, the real thing under NCCL on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _torchrun(n, port, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("mode,mesh", [(0, "cartesian"), (1, "cartesian"), (1, "hex")])
def test_sharding_host_logic_gloo(mode, mesh):
    r = _torchrun(2, 29611 + mode + (4 if mesh == "hex" else 0), "--backend", "gloo", "--shard-mode", str(mode),
                  "--mesh", mesh)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode,mesh", [(0, "cartesian"), (1, "cartesian"), (1, "hex")])
def test_sharded_solve_matches_single_gpu(mode, mesh):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, 29621 + mode + (4 if mesh == "hex" else 0), "--backend", "nccl", "--shard-mode", str(mode),
                  "--mesh", mesh)
    assert r.returncode == 0 and "NCCL_OK" in r.stdout, r.stdout + r.stderr
