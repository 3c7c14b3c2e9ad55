"""Parity with the oracle AT THE SIZES THAT ARE BENCHMARKED (BASELINE.json configs 2, 4 and 5 at 1/16 scale): forward and
backward of the whole stack through the module API -> C ABI, against the reference-pinned CPU oracle on the same seeded
synthetic inputs that bench.py times.  Outputs are judged on the UPDATE (x' - x, Z' - Z), see tests/gpu_util.py.

The oracle runs in fp64 where the host has the memory for it (the reference keeps ~6 KB of activations per edge and
layer in fp32); config 5 at 1/16 scale (62 500 nodes, ~1.9e6 edges, C=8, 4 layers) runs the oracle in fp32, whose own
distance from fp64 (~1e-6) is far below the stated TF32 tolerances and is added to the fp32-mode bound."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fastegnn_oracle as orc
from tests.gpu_util import build_gpu_model, precision, rel_err, update_err

pytestmark = pytest.mark.gpu


def _workload(name):
    import bench
    if name == "water3d":            # config 4: the headline workload of bench.py, as bench.py builds it
        return bench.make_cloud(8000, 25.0, 3, 0, [0, -1, 0]), torch.float64
    if name == "nbody100":           # config 2 shape: 10 graphs x 100 particles, 4 950 edges each (cutoff 0.5)
        return bench.make_nbody(100, 10, 0.5, 3, 0), torch.float64
    if name == "large_16th":         # config 5 at 1/16 scale: 62 500 nodes, mean degree 30, C=8
        return bench.make_cloud(62500, 30.0, 8, 0, None), torch.float32
    raise KeyError(name)


def _oracle(cfg, params, data, wx, wz, dtype):
    p = {k: v.to(dtype).requires_grad_(True) for k, v in params.items()}
    f = lambda k: data[k].detach().clone().to(dtype)
    x0, lm = f("loc_0").requires_grad_(True), f("loc_mean").requires_grad_(True)
    x, Z = orc.fastegnn_forward(p, cfg, f("node_feat"), x0, f("vel_0"), data["edge_index"], data["batch"], lm,
                                f("edge_attr"))
    ((x * wx.to(dtype)).sum() + (Z * wz.to(dtype)).sum()).backward()
    return dict(x=x.detach(), Z=Z.detach(), gx0=x0.grad, glm=lm.grad, gp={k: v.grad for k, v in p.items()})


@pytest.mark.parametrize("gain", [1.0, 100.0])
@pytest.mark.parametrize("name", ["water3d", "nbody100", "large_16th"])
def test_benchmarked_configs_against_oracle(name, gain):
    """gain 1 = the reference's own initialisation (what bench.py times; coordinate heads xavier(gain=1e-3), so the
    update is ~1e-3 of the coordinates); gain 100 = the same weights with the coordinate heads at natural magnitude, so
    that the edge path carries the update."""
    if name == "large_16th" and gain != 1.0:
        pytest.skip("one oracle pass at 1.9e6 edges is enough (minutes of CPU)")
    data, odt = _workload(name)
    C = data["C"]
    cfg = orc.OracleConfig(node_feat_nf=2, edge_attr_nf=2, hidden_nf=64, virtual_channels=C, n_layers=4,
                           gravity=data["gravity"])
    params = orc.make_params(cfg, 0)
    if gain != 1.0:
        orc.rescale_coord_heads(params, gain)
    N, B = data["loc_0"].size(0), data["n_graphs"]
    g = torch.Generator().manual_seed(17)
    wx, wz = torch.randn(N, 3, generator=g) / N, torch.randn(B, 3, C, generator=g) / B
    torch.set_num_threads(os.cpu_count() or 1)
    ref = _oracle(cfg, params, data, wx, wz, odt)
    slack = 0.0 if odt == torch.float64 else 4e-6          # fp32 oracle: its own rounding
    dev = "cuda:0"
    lines = []
    for prec in ("tf32", "fp32"):
        with precision(prec) as tol:
            m = build_gpu_model(cfg, params, dev)
            x0 = data["loc_0"].to(dev).requires_grad_(True)
            lm = data["loc_mean"].to(dev).requires_grad_(True)
            x, Z = m(node_feat=data["node_feat"].to(dev), node_loc=x0, node_vel=data["vel_0"].to(dev),
                     edge_index=data["edge_index"].to(dev), data_batch=data["batch"].to(dev), loc_mean=lm,
                     edge_attr=data["edge_attr"].to(dev))
            ((x * wx.to(dev)).sum() + (Z * wz.to(dev)).sum()).backward()
            torch.cuda.synchronize()
        ex = update_err(x.detach().cpu(), ref["x"], data["loc_0"])
        ez = update_err(Z.detach().cpu(), ref["Z"], data["loc_mean"])
        egx, egl = rel_err(x0.grad.cpu(), ref["gx0"]), rel_err(lm.grad.cpu(), ref["glm"])
        worst_w, worst_k, worst_b, over = 0.0, "", 0.0, []
        for k, p in m.named_parameters():
            if ref["gp"][k] is None:
                assert p.grad is None, k
                continue
            e = rel_err(p.grad.cpu(), ref["gp"][k])
            if tol.for_param(k) != tol.gw:               # the two cancellation-dominated head biases (tests/gpu_util.py)
                worst_b = max(worst_b, e)
            elif e > worst_w:
                worst_w, worst_k = e, k
            if e > tol.for_param(k) + 10 * slack:
                over.append((k, e))
        lines.append(f"{name} gain={gain:g} [{prec}] N={N} E={data['edge_index'].size(1)} C={C} oracle={odt}: "
                     f"x'-x {ex:.2e}  Z'-Z {ez:.2e}  g_x0 {egx:.2e}  g_loc_mean {egl:.2e}  "
                     f"weight grads {worst_w:.2e} ({worst_k})  virtual-head biases {worst_b:.2e}   "
                     f"stated {tol.out:g} / {tol.gin:g} / {tol.gw:g} / {tol.gb:g}")
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/parity_fullsize.txt", "a") as f:
            f.write(lines[-1] + "\n")
        assert ex < tol.out + slack and ez < tol.out + slack, lines[-1]
        assert egx < tol.gin + 10 * slack and egl < tol.gin + 10 * slack, lines[-1]
        assert not over, (over, lines[-1])
