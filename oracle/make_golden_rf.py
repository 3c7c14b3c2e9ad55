"""Generate tests/golden/rf_*.{npz,json} from the UNMODIFIED reference models/FastRF.py.

Run in the build container only (needs /root/reference):   python oracle/make_golden_rf.py
Same recipe as oracle/make_golden.py (whose helpers it reuses): stand-in for torch_geometric.nn.global_mean_pool,
seeded construction, forward + backward of a fixed linear functional, outputs / input gradients / digests of every
parameter gradient (full gradients for the narrow H=16 case).  Test infrastructure; never imported by the product.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import fastegnn_oracle as orc  # noqa: E402
from oracle import fastrf_oracle as rfo  # noqa: E402
from oracle import make_golden as mg  # noqa: E402

CASES = [
    dict(name="rf_c3_gravity", seed=21, data_seed=201, hidden_nf=64, virtual_channels=3, n_layers=4, node_feat_nf=2,
         edge_attr_nf=2, graph_sizes=[40], edges_per_graph=300, gain=1000.0, gravity=[0, -1, 0]),
    dict(name="rf_c3_batch3", seed=22, data_seed=202, hidden_nf=64, virtual_channels=3, n_layers=4, node_feat_nf=2,
         edge_attr_nf=2, graph_sizes=[7, 12, 5], edges_per_graph=30, gain=1000.0),
    dict(name="rf_c2_flags", seed=23, data_seed=203, hidden_nf=64, virtual_channels=2, n_layers=2, node_feat_nf=2,
         edge_attr_nf=2, graph_sizes=[10, 6], edges_per_graph=25, gain=1000.0, attention=True, normalize=True, tanh=True),
    dict(name="rf_h16_full_grads", seed=24, data_seed=204, hidden_nf=16, virtual_channels=2, n_layers=3, node_feat_nf=2,
         edge_attr_nf=2, graph_sizes=[6, 8], edges_per_graph=20, gain=1000.0, gravity=[0, -1, 0]),
]


def main():
    mg._install_pyg_standin()
    spec = importlib.util.spec_from_file_location("ref_FastRF", os.path.join(mg.REF, "models", "FastRF.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for case in CASES:
        torch.manual_seed(case["seed"])
        model = ref.FastRF(node_feat_nf=case["node_feat_nf"], node_attr_nf=0, edge_attr_nf=case["edge_attr_nf"],
                           hidden_nf=case["hidden_nf"], virtual_channels=case["virtual_channels"], device="cpu",
                           n_layers=case["n_layers"], residual=True, attention=case.get("attention", False),
                           normalize=case.get("normalize", False), tanh=case.get("tanh", False),
                           gravity=case.get("gravity"))
        cfg = orc.OracleConfig(node_feat_nf=case["node_feat_nf"], edge_attr_nf=case["edge_attr_nf"],
                               hidden_nf=case["hidden_nf"], virtual_channels=case["virtual_channels"],
                               n_layers=case["n_layers"], attention=case.get("attention", False),
                               normalize=case.get("normalize", False), tanh=case.get("tanh", False),
                               gravity=case.get("gravity"))
        replay = rfo.make_params(cfg, case["seed"])
        sd = model.state_dict()
        assert set(sd.keys()) == set(replay.keys()), (set(sd) ^ set(replay))
        for k in sd:
            assert torch.equal(sd[k], replay[k]), f"RNG replay mismatch at {k}"
        if case["gain"] != 1.0:
            with torch.no_grad():
                for k, p in model.named_parameters():
                    if k.endswith(".2.weight") and ("coord_mlp_r" in k or "coord_mlp_v_virtual" in k):
                        p.mul_(case["gain"])
        inp = mg.make_case_inputs(case)
        leaf = {k: inp[k].clone().requires_grad_(True) for k in ("node_loc", "loc_mean", "node_feat")}
        x, Z = model(node_feat=leaf["node_feat"], node_loc=leaf["node_loc"], node_vel=inp["node_vel"],
                     edge_index=inp["edge_index"], data_batch=inp["data_batch"], loc_mean=leaf["loc_mean"],
                     edge_attr=inp["edge_attr"], node_attr=None)
        loss = (x * inp["wx"]).sum() + (Z * inp["wz"]).sum()
        loss.backward()
        arrays = {f"in_{k}": v.numpy() for k, v in inp.items()}
        arrays["out_x"], arrays["out_Z"] = x.detach().numpy(), Z.detach().numpy()
        for k, t in leaf.items():
            arrays[f"gin_{k}"] = t.grad.numpy()
        meta = dict(case=case, loss=float(loss), keys=list(sd.keys()), param_sha256={}, grad_digest={}, grad_none=[])
        for k, p in model.named_parameters():
            meta["param_sha256"][k] = hashlib.sha256(p.detach().numpy().tobytes()).hexdigest()
            if p.grad is None:
                meta["grad_none"].append(k)
            else:
                meta["grad_digest"][k] = mg._digest(p.grad)
                if case["hidden_nf"] <= 16:
                    arrays[f"gp_{k}"] = p.grad.numpy()
        np.savez_compressed(os.path.join(mg.OUT, f"{case['name']}.npz"), **arrays)
        with open(os.path.join(mg.OUT, f"{case['name']}.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(f"{case['name']}: N={x.size(0)} E={inp['edge_index'].size(1)} loss={float(loss):.6f} "
              f"grad_none={len(meta['grad_none'])}")


if __name__ == "__main__":
    main()
