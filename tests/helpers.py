"""Shared test helpers: golden-fixture loading and oracle plumbing (test infrastructure)."""
import hashlib
import json
import os

import numpy as np
import torch

from oracle import fastegnn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_CASES = ["c3_batch3", "c3_gravity", "c8_two_layers", "c1_flags", "equiv_shape_default_init", "h16_full_grads"]
H64_CASES = [c for c in MODEL_CASES if c != "h16_full_grads"]


def load_case(name):
    meta = json.load(open(os.path.join(GOLDEN, f"{name}.json")))
    arr = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    return meta, arr


def case_config(case) -> orc.OracleConfig:
    return orc.OracleConfig(node_feat_nf=case["node_feat_nf"], edge_attr_nf=case["edge_attr_nf"],
                            hidden_nf=case["hidden_nf"], virtual_channels=case["virtual_channels"],
                            n_layers=case["n_layers"], attention=case.get("attention", False),
                            normalize=case.get("normalize", False), tanh=case.get("tanh", False),
                            gravity=case.get("gravity"))


def case_params(case, dtype=torch.float32):
    cfg = case_config(case)
    params = orc.make_params(cfg, case["seed"])
    if case["gain"] != 1.0:
        orc.rescale_coord_heads(params, case["gain"])
    return cfg, {k: v.to(dtype) for k, v in params.items()}


def case_inputs(arr, dtype=torch.float32):
    out = {}
    for k in ("node_feat", "node_loc", "node_vel", "loc_mean", "edge_attr", "wx", "wz"):
        out[k] = torch.from_numpy(arr[f"in_{k}"]).to(dtype)
    out["edge_index"] = torch.from_numpy(arr["in_edge_index"])
    out["data_batch"] = torch.from_numpy(arr["in_data_batch"])
    return out


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().numpy().tobytes()).hexdigest()


def oracle_run(cfg, params, inp, want_grads=True):
    """Forward (+ backward of the fixture's linear functional) through the oracle."""
    p = {k: v.clone().requires_grad_(want_grads) for k, v in params.items()}
    leaf = {k: inp[k].clone().requires_grad_(want_grads) for k in ("node_loc", "node_vel", "loc_mean", "node_feat")}
    x, Z = orc.fastegnn_forward(p, cfg, leaf["node_feat"], leaf["node_loc"], leaf["node_vel"], inp["edge_index"],
                                inp["data_batch"], leaf["loc_mean"], inp["edge_attr"])
    res = dict(x=x.detach(), Z=Z.detach())
    if want_grads:
        loss = (x * inp["wx"]).sum() + (Z * inp["wz"]).sum()
        loss.backward()
        res["loss"] = float(loss.detach())
        res["gin"] = {k: t.grad for k, t in leaf.items()}
        res["gp"] = {k: t.grad for k, t in p.items()}
    return res


LAYER_CASES = ["layer_sum", "layer_mean_gravity"]


def load_layer_case(name):
    """Golden vectors of ONE reference layer (oracle/make_golden_layer.py): returns (cfg, params keyed 'gcl_0.*', arrays)."""
    arr = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    grav = arr["gravity"].tolist() or None
    C = int(arr["in_S"].shape[2])
    cfg = orc.OracleConfig(node_feat_nf=2, edge_attr_nf=int(arr["in_edge_attr"].shape[1]), hidden_nf=64,
                           virtual_channels=C, n_layers=1, gravity=grav, coords_agg=str(arr["coords_agg"]))
    params = {"gcl_0." + k[2:]: torch.from_numpy(v) for k, v in arr.items() if k.startswith("p_")}
    return cfg, params, arr
