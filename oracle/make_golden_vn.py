"""Golden vectors of the reference's VNEGNN sibling (models/VNEGNN.py), produced by the UNMODIFIED file.
Run in the build container only (needs /root/reference):   python oracle/make_golden_vn.py
Stand-ins: torch_geometric.nn.global_mean_pool (as in make_golden.py) and an empty torch_scatter module (the file imports
scatter_add at :6 and never calls it).  Test infrastructure; never imported by the product."""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import OUT, REF, _install_pyg_standin, make_case_inputs  # noqa: E402

CASES = [
    dict(name="vn_c3_batch2", seed=41, data_seed=141, virtual_channels=3, n_layers=2, graph_sizes=[9, 7], edges_per_graph=30,
         node_feat_nf=2, edge_attr_nf=2, normalize=False, tanh=False),
    dict(name="vn_c2_flags", seed=42, data_seed=142, virtual_channels=2, n_layers=1, graph_sizes=[12], edges_per_graph=40,
         node_feat_nf=2, edge_attr_nf=2, normalize=True, tanh=True),
]


def main():
    _install_pyg_standin()
    ts = types.ModuleType("torch_scatter")
    ts.scatter_add = None
    sys.modules["torch_scatter"] = ts
    spec = importlib.util.spec_from_file_location("ref_VNEGNN", os.path.join(REF, "models", "VNEGNN.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for case in CASES:
        torch.manual_seed(case["seed"])
        m = ref.VNEGNN(node_feat_nf=case["node_feat_nf"], node_attr_nf=0, edge_attr_nf=case["edge_attr_nf"], hidden_nf=64,
                       virtual_channels=case["virtual_channels"], device="cpu", n_layers=case["n_layers"],
                       normalize=case["normalize"], tanh=case["tanh"])
        sha = {k: hashlib.sha256(v.detach().numpy().tobytes()).hexdigest() for k, v in m.state_dict().items()}
        with torch.no_grad():                       # natural magnitude on the coordinate path (xavier gain 1e-3 otherwise)
            for n, p in m.named_parameters():
                if n.endswith("coord_mlp.2.weight"):
                    p.mul_(300.0)
        inp = make_case_inputs(case)
        x = inp["node_loc"].clone().requires_grad_(True)
        Z = inp["loc_mean"].clone().requires_grad_(True)
        xo, Zo = m(node_feat=inp["node_feat"], node_loc=x, edge_index=inp["edge_index"], data_batch=inp["data_batch"],
                   virtual_node_loc=Z, edge_attr=inp["edge_attr"], node_attr=None)
        ((xo * inp["wx"]).sum() + (Zo * inp["wz"]).sum()).backward()
        out = {f"p_{k}": v.detach().numpy() for k, v in m.state_dict().items()}
        out.update({f"gp_{k}": p.grad.numpy() for k, p in m.named_parameters() if p.grad is not None})
        out["grad_none"] = np.array(sorted(k for k, p in m.named_parameters() if p.grad is None))
        out.update({f"sha_{k}": np.array(v) for k, v in sha.items()})
        out.update(out_x=xo.detach().numpy(), out_Z=Zo.detach().numpy(), g_x=x.grad.numpy(), g_Z=Z.grad.numpy(),
                   seed=np.array(case["seed"]), n_layers=np.array(case["n_layers"]),
                   normalize=np.array(case["normalize"]), tanh=np.array(case["tanh"]))
        out.update({f"in_{k}": v.numpy() for k, v in inp.items()})
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **out)
        print("wrote", case["name"], "x' - x max", float((xo.detach() - x.detach()).abs().max()),
              "Z' - Z max", float((Zo.detach() - Z.detach()).abs().max()))


if __name__ == "__main__":
    main()
