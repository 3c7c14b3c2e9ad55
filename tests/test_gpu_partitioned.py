"""Partitioned multi-GPU path vs the single-GPU path (needs >= 2 GPUs; run with gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


PORTS = {"nccl": 0, "p2p": 20, "fused": 40}


@pytest.mark.parametrize("plan", ["host", "device"])
@pytest.mark.parametrize("halo", ["nccl", "p2p", "fused"])
@pytest.mark.parametrize("C", [3, 8])
def test_partitioned_matches_single_gpu(C, halo, plan):
    """halo: NCCL all-to-all / peer-memory kernels + symmetric-memory barrier / fused payload+signal kernels with the
    deterministic reverse halo and one-shot all-reduces.  plan: numpy slab plan over the global edge list / every rank
    builds its own slab graph on its device (DeviceSlabPlan)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    if plan == "device" and halo == "p2p":
        pytest.skip("covered by the other two transports")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, CHECK_C=str(C), CHECK_HALO=halo, CHECK_PLAN=plan)
    port = 29500 + C + PORTS[halo] + (100 if plan == "device" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
