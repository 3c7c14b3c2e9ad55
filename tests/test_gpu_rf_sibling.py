"""GPU: the FastRF sibling (models/FastRF.py) on the FastEGNN kernels (FEGNN_F_RF) against the reference's golden
vectors and the fp64 oracle."""
import numpy as np
import pytest
import torch

from tests.gpu_util import TOLERANCES, make_graph_case, precision, rel_err
from tests.helpers import case_inputs, load_case
from tests.test_fastrf_cpu import RF_H64, rf_case_params, rf_oracle_run

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build_rf(cfg, params):
    from fastegnn_b200 import FastRF
    m = FastRF(node_feat_nf=cfg.node_feat_nf, node_attr_nf=0, edge_attr_nf=cfg.edge_attr_nf, hidden_nf=cfg.hidden_nf,
               virtual_channels=cfg.virtual_channels, device=DEV, n_layers=cfg.n_layers, attention=cfg.attention,
               normalize=cfg.normalize, tanh=cfg.tanh, gravity=cfg.gravity)
    m.load_state_dict({k: v.to(DEV) for k, v in params.items()})
    return m


def gpu_rf_run(cfg, params, inp):
    m = build_rf(cfg, params)
    g = {k: v.to(DEV) for k, v in inp.items()}
    leaf = {k: g[k].clone().requires_grad_(True) for k in ("node_loc", "loc_mean", "node_feat")}
    x, Z = m(node_feat=leaf["node_feat"], node_loc=leaf["node_loc"], node_vel=g["node_vel"], edge_index=g["edge_index"],
             data_batch=g["data_batch"], loc_mean=leaf["loc_mean"], edge_attr=g["edge_attr"])
    ((x * g["wx"]).sum() + (Z * g["wz"]).sum()).backward()
    torch.cuda.synchronize()
    return dict(x=x.detach().cpu(), Z=Z.detach().cpu(), gin={k: t.grad.cpu() for k, t in leaf.items()},
                gp={k: (None if p.grad is None else p.grad.cpu()) for k, p in m.named_parameters()})


def compare(cfg, params, inp, res, tol_out, tol_grad, prec):
    p64 = {k: v.double() for k, v in params.items()}
    i64 = {k: (v.double() if v.is_floating_point() else v) for k, v in inp.items()}
    r64 = rf_oracle_run(cfg, p64, i64)
    assert rel_err(res["x"], r64["x"]) < tol_out
    assert rel_err(res["Z"], r64["Z"]) < tol_out
    for k in ("node_loc", "loc_mean", "node_feat"):
        assert rel_err(res["gin"][k], r64["gin"][k]) < tol_grad, k
    for k, g64 in r64["gp"].items():
        assert res["gp"][k] is not None, k
        assert rel_err(res["gp"][k], g64) < (tol_grad if prec == "fp32" else 5 * tol_grad), k


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("name", RF_H64)
def test_rf_golden_vectors_from_reference(name, prec):
    meta, arr = load_case(name)
    cfg, params = rf_case_params(meta["case"])
    inp = case_inputs(arr)
    with precision(prec) as (tol_out, tol_grad):
        res = gpu_rf_run(cfg, params, inp)
    assert rel_err(res["x"], torch.from_numpy(arr["out_x"])) < tol_out
    assert rel_err(res["Z"], torch.from_numpy(arr["out_Z"])) < tol_out
    compare(cfg, params, inp, res, tol_out, tol_grad, prec)
    if not cfg.normalize:             # golden gradients of normalize=True carry self-loop cancellation noise
        for k, dig in meta["grad_digest"].items():
            g = res["gp"][k].double().flatten()
            np.testing.assert_allclose(g[dig["idx"]].numpy(), np.array(dig["val"]), rtol=0,
                                       atol=tol_grad * (dig["l2"] + 1e-30), err_msg=k)


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("kw", [dict(seed=31, sizes=[300, 211, 190], deg=12, C=3),
                                dict(seed=32, sizes=[500], deg=20, C=3, gravity=[0, -1, 0], heavy_row=400),
                                dict(seed=33, sizes=[5] * 60, deg=2, C=8, L=2)])
def test_rf_seeded_batches_against_oracle(kw, prec):
    from oracle import fastegnn_oracle as orc
    from oracle import fastrf_oracle as rfo
    kw = dict(kw)
    seed = kw.pop("seed")
    cfg, _, inp = make_graph_case(seed, kw.pop("sizes"), kw.pop("deg"), kw.pop("C"), **kw)
    params = rfo.make_params(cfg, seed + 1000)
    orc.rescale_coord_heads(params, 1000.0)
    with precision(prec) as (tol_out, tol_grad):
        res = gpu_rf_run(cfg, params, inp)
    compare(cfg, params, inp, res, tol_out, tol_grad, prec)
