#!/usr/bin/env python
"""CPU numerics study for DESIGN.md section 6 item 2: what happens to the weight gradients of the fused edge backward
if the operand tiles of its weight-gradient GEMMs (m, a1, g3, g2, gz1 -- 128 edges x 64 columns each) are stored as
  tf32      10-bit mantissa, fp32 exponent          (what edge_bwd_tc2_kernel does today)
  bf16      8-bit mantissa                          (halves the tile context, no scaling needed)
  fp16      10-bit mantissa, 5-bit exponent, unscaled
  fp16s     fp16 after dividing each 128 x 64 tile by its max-abs (one fp32 scale per tile, folded back after the GEMM)
  fp16g     gradient tiles fp16 with ONE power-of-two scale per launch and tensor, activation tiles plain fp16
  mixed     gradient tiles bf16, activation tiles fp16
Everything else of the 4-layer backward stays fp64 (oracle/staged.py), so the numbers isolate the operand storage.
Reported: worst relative error (of the tensor's max) of the edge-phase weight gradients over the layers, per format.

    python tools/wgrad_quant_study.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import staged  # noqa: E402
from tests.gpu_util import make_graph_case  # noqa: E402

TILE = 128


def q_mantissa(x, bits):
    """round-to-nearest-even to `bits` explicit mantissa bits (fp32 exponent range)."""
    x32 = x.float()
    i = x32.view(torch.int32)
    drop = 23 - bits
    bias = (1 << (drop - 1)) - 1 + ((i >> drop) & 1)
    return ((i + bias) & ~((1 << drop) - 1)).view(torch.float32).double()


def q_tile(x, fmt):
    if fmt == "fp64":
        return x
    if fmt == "tf32":
        return q_mantissa(x, 10)
    if fmt == "bf16":
        return x.to(torch.bfloat16).double()
    if fmt == "fp16":
        return x.to(torch.float16).double()
    if fmt == "fp16s":
        E = x.size(0)
        out = torch.empty_like(x)
        for t0 in range(0, E, TILE):
            blk = x[t0:t0 + TILE]
            s = blk.abs().max().clamp(min=1e-300)
            out[t0:t0 + TILE] = (blk / s).to(torch.float16).double() * s
        return out
    if fmt == "fp16g":          # ONE power-of-two scale per launch: the tensor's max lands at 2^8 (a bound pre-pass would be looser)
        s = 2.0 ** (torch.floor(torch.log2(x.abs().max().clamp(min=1e-300))) - 8)
        return (x / s).to(torch.float16).double() * s
    raise ValueError(fmt)


def edge_bwd_quant(fmt):
    ref = staged.edge_bwd

    def f(w, g, fl, P, Q, x, ea, gm, gt):
        out = ref(w, g, fl, P, Q, x, ea, gm, gt)
        r = staged._edge_recompute(w, g, fl, P, Q, x, ea)
        row = g.row
        gte = gt[row]
        gs = (r["dn"] * gte).sum(1)
        gz3 = gs[:, None] * w.w4 * staged.dsilu(r["z3"])
        gmm = gm[row] + gz3 @ w.W3
        gz2 = gmm * staged.dsilu(r["z2"])
        gz1 = (gz2 @ w.W2) * staged.dsilu(r["z1"])
        # "mixed": gradient tiles bf16 (no scale needed), activation tiles fp16; "fp16g": gradients with one scale per
        # launch, activations plain fp16
        qg = lambda t: q_tile(t, "bf16" if fmt == "mixed" else fmt)
        qa = lambda t: q_tile(t, "fp16" if fmt in ("mixed", "fp16g") else fmt)
        wg = dict(out["wg"])
        ones_q_ea = torch.cat([torch.ones_like(r["q"])[:, None], r["q"][:, None], ea], dim=1)
        aux = qa(ones_q_ea)                                 # the aux tile (1, q, ea...)
        wg["W3"], wg["b3"] = qg(gz3).T @ qa(r["m"]), qg(gz3).T @ aux[:, 0]
        wg["W2"], wg["b2"] = qg(gz2).T @ qa(r["a1"]), qg(gz2).T @ aux[:, 0]
        wg["wq"], wg["Wa"] = qg(gz1).T @ aux[:, 1], qg(gz1).T @ aux[:, 2:]
        out["wg"] = wg
        return out
    return f


def run(case_kw, fmt):
    cfg, params, inp = make_graph_case(**case_kw)
    p64 = {k: v.double() for k, v in params.items()}
    i64 = {k: (v.double() if v.is_floating_point() else v) for k, v in inp.items()}
    fl = staged.Flags(cfg.attention, cfg.normalize, cfg.tanh, cfg.gravity, cfg.eps)
    sm = staged.StagedModel(p64, cfg.hidden_nf, cfg.virtual_channels, cfg.edge_attr_nf, cfg.n_layers, fl)
    keep = staged.edge_bwd
    staged.edge_bwd = edge_bwd_quant(fmt)
    try:
        sm.forward(i64["node_feat"], i64["node_loc"], i64["node_vel"], i64["edge_index"], i64["data_batch"],
                   i64["loc_mean"], i64["edge_attr"])
        grads, _ = sm.backward(i64["wx"], i64["wz"])
    finally:
        staged.edge_bwd = keep
    return grads


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cases = {
        "multi_tile_c3 (coordinate heads x1000)": dict(seed=1, sizes=[300, 211, 190], deg=12, C=3),
        "default_gain (heads at the reference's 1e-3 init)": dict(seed=8, sizes=[128, 128], deg=8, C=3, gain=1.0),
        "gravity_heavy_row": dict(seed=2, sizes=[500], deg=20, C=3, gravity=[0, -1, 0], heavy_row=400),
    }
    keys = ("edge_mlp.0.weight", "edge_mlp.2.weight", "edge_mlp.2.bias", "coord_mlp_r.0.weight", "coord_mlp_r.0.bias")
    for name, kw in cases.items():
        ref = run(kw, "fp64")
        print(f"== {name}")
        for fmt in ("tf32", "fp16s", "fp16g", "mixed", "bf16", "fp16"):
            g = run(kw, fmt)
            worst, where = 0.0, ""
            for k, v in ref.items():
                if not k.endswith(keys):
                    continue
                e = float((g[k] - v).abs().max() / (v.abs().max() + 1e-300))
                if e > worst:
                    worst, where = e, k
            print(f"   {fmt:6s} worst rel. error of the edge-phase weight gradients {worst:.2e}   ({where})")


if __name__ == "__main__":
    main()
