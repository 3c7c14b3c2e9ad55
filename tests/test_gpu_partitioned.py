"""Partitioned multi-GPU path vs the single-GPU path (needs >= 2 GPUs; run with gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("halo", ["nccl", "p2p"])
@pytest.mark.parametrize("C", [3, 8])
def test_partitioned_matches_single_gpu(C, halo):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, CHECK_C=str(C), CHECK_HALO=halo)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + C + (20 if halo == "p2p" else 0)), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
