"""Golden vectors of ONE reference layer (`E_GCL_vel`, models/FastEGNN.py:6-223) for both `coords_agg` values, and of
the module-level helpers `unsorted_segment_sum` / `unsorted_segment_mean` (:279-294), produced by the UNMODIFIED
reference file.  Run in the build container only (needs /root/reference):

    python oracle/make_golden_layer.py

`FastEGNN` never passes `coords_agg` (:261), so the model-level goldens of make_golden.py cannot pin the 'sum' branch
(:124-125); this script calls the layer class directly.  Test infrastructure; never imported by the product.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import OUT, _import_reference_model, make_case_inputs  # noqa: E402

CASES = [
    dict(name="layer_sum", coords_agg="sum", seed=31, data_seed=131, virtual_channels=3, graph_sizes=[11, 8],
         edges_per_graph=40, node_feat_nf=2, edge_attr_nf=2, gravity=None),
    dict(name="layer_mean_gravity", coords_agg="mean", seed=32, data_seed=132, virtual_channels=2, graph_sizes=[14],
         edges_per_graph=60, node_feat_nf=2, edge_attr_nf=2, gravity=[0, -1, 0]),
]


def main():
    ref = _import_reference_model()
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        torch.manual_seed(case["seed"])
        grav = None if case["gravity"] is None else torch.tensor(case["gravity"])
        layer = ref.E_GCL_vel(64, 64, 0, case["edge_attr_nf"], 64, virtual_channels=case["virtual_channels"],
                              coords_agg=case["coords_agg"], gravity=grav)
        with torch.no_grad():                       # natural magnitude on the coordinate path (xavier gain 1e-3 otherwise)
            for n, p in layer.named_parameters():
                if n.endswith(".2.weight") and ("coord_mlp_r" in n or "coord_mlp_v_virtual" in n):
                    p.mul_(300.0)
        inp = make_case_inputs(case)
        g = torch.Generator().manual_seed(case["data_seed"] + 7)
        N, B, C = inp["node_loc"].size(0), len(case["graph_sizes"]), case["virtual_channels"]
        h = torch.randn(N, 64, generator=g).requires_grad_(True)
        S = torch.randn(B, 64, C, generator=g).requires_grad_(True)
        x = inp["node_loc"].clone().requires_grad_(True)
        Z = inp["loc_mean"].clone().requires_grad_(True)
        wh, wS = torch.randn(N, 64, generator=g), torch.randn(B, 64, C, generator=g)
        ho, xo, So, Zo = layer(h, inp["edge_index"], x, inp["node_vel"], Z, S, inp["data_batch"],
                               edge_attr=inp["edge_attr"])
        ((ho * wh).sum() + (xo * inp["wx"]).sum() + (So * wS).sum() + (Zo * inp["wz"]).sum()).backward()
        out = {f"p_{k}": v.detach().numpy() for k, v in layer.state_dict().items()}
        out.update({f"gp_{k}": p.grad.numpy() for k, p in layer.named_parameters()})
        out.update(in_h=h.detach().numpy(), in_S=S.detach().numpy(), wh=wh.numpy(), wS=wS.numpy(),
                   out_h=ho.detach().numpy(), out_x=xo.detach().numpy(), out_S=So.detach().numpy(),
                   out_Z=Zo.detach().numpy(), g_h=h.grad.numpy(), g_x=x.grad.numpy(), g_S=S.grad.numpy(),
                   g_Z=Z.grad.numpy(), coords_agg=np.array(case["coords_agg"]),
                   gravity=np.array(case["gravity"] if case["gravity"] is not None else [], dtype=np.float32))
        out.update({f"in_{k}": v.numpy() for k, v in inp.items()})
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **out)
        print("wrote", case["name"], "x' - x max", float((xo - x).abs().max()))
    # module-level helpers (:279-294)
    g = torch.Generator().manual_seed(5)
    data = torch.randn(57, 3, generator=g)
    ids = torch.randint(0, 9, (57,), generator=g)          # segments 9..11 stay empty
    np.savez_compressed(os.path.join(OUT, "segment_helpers.npz"), data=data.numpy(), ids=ids.numpy(), num=np.array(12),
                        sum=ref.unsorted_segment_sum(data, ids, 12).numpy(),
                        mean=ref.unsorted_segment_mean(data, ids, 12).numpy())
    print("wrote segment_helpers")


if __name__ == "__main__":
    main()
