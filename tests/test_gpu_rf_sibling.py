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
    from tests.gpu_util import TOLERANCES, update_err
    tol = TOLERANCES[prec]
    p64 = {k: v.double() for k, v in params.items()}
    i64 = {k: (v.double() if v.is_floating_point() else v) for k, v in inp.items()}
    r64 = rf_oracle_run(cfg, p64, i64)
    assert update_err(res["x"], r64["x"], inp["node_loc"]) < tol.out          # judged on the update x' - x / Z' - Z
    assert update_err(res["Z"], r64["Z"], inp["loc_mean"]) < tol.out
    for k in ("node_loc", "loc_mean", "node_feat"):
        assert rel_err(res["gin"][k], r64["gin"][k]) < tol.gin, k
    for k, g64 in r64["gp"].items():
        assert res["gp"][k] is not None, k
        assert rel_err(res["gp"][k], g64) < tol.for_param(k), k


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("name", RF_H64)
def test_rf_golden_vectors_from_reference(name, prec):
    meta, arr = load_case(name)
    cfg, params = rf_case_params(meta["case"])
    inp = case_inputs(arr)
    with precision(prec) as (tol_out, tol_grad):
        res = gpu_rf_run(cfg, params, inp)
    from tests.gpu_util import TOLERANCES, update_err
    tol = TOLERANCES[prec]
    assert update_err(res["x"], torch.from_numpy(arr["out_x"]), inp["node_loc"]) < tol.out + 4e-6
    assert update_err(res["Z"], torch.from_numpy(arr["out_Z"]), inp["loc_mean"]) < tol.out + 4e-6
    compare(cfg, params, inp, res, tol_out, tol_grad, prec)
    if not cfg.normalize:             # golden gradients of normalize=True carry self-loop cancellation noise
        for k, dig in meta["grad_digest"].items():
            g = res["gp"][k].double().flatten()
            np.testing.assert_allclose(g[dig["idx"]].numpy(), np.array(dig["val"]), rtol=0,
                                       atol=tol.for_param(k) * (dig["l2"] + 1e-30), err_msg=k)


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


@pytest.mark.parametrize("model_name", ["fastrf", "fastegnn"])
def test_protein_shape_graph_built_on_device(model_name):
    """Config 3 shape end to end on the device: 3 frames x 855 atoms, 10 A contact graph with the shortest 50% kept
    (datasets/protein/dataset.py:146-156,208-213) from CsrGraph.from_radius, consumed as a prebuilt graph by FastRF
    (main_protein.py:114) and FastEGNN; compared with the oracle on the exported int64 edge list."""
    from bench import make_protein
    from fastegnn_b200 import CsrGraph, FastEGNN, FastRF
    from oracle import fastegnn_oracle as orc
    from oracle import fastrf_oracle as rfo
    from oracle import radius_graph_oracle as rgo
    data = make_protein(855, 3, 0.5, 3, seed=4)
    x = data["loc_0"]
    g = CsrGraph.from_radius(x.to(DEV), data["batch"].to(DEV), 3, 10.0, 0.5, 2)
    o = rgo.radius_graph_csr(x.numpy(), np.arange(4) * 855, 10.0, 0.5)
    np.testing.assert_array_equal(g.row.cpu().numpy(), o["row"])
    np.testing.assert_array_equal(g.col.cpu().numpy(), o["col"])
    cfg = orc.OracleConfig(node_feat_nf=2, edge_attr_nf=2, hidden_nf=64, virtual_channels=3, n_layers=4)
    rf = model_name == "fastrf"
    params = (rfo if rf else orc).make_params(cfg, 77)
    orc.rescale_coord_heads(params, 1000.0)
    m = (FastRF if rf else FastEGNN)(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=3,
                                     device=DEV, n_layers=4)
    m.load_state_dict({k: v.to(DEV) for k, v in params.items()})
    ei, ea = g.edge_index().cpu(), g.edge_attr.cpu()
    p64 = {k: v.double() for k, v in params.items()}
    fwd = rfo.fastrf_forward if rf else orc.fastegnn_forward
    x64, Z64 = fwd(p64, cfg, data["node_feat"].double(), x.double(), data["vel_0"].double(), ei, data["batch"],
                   data["loc_mean"].double(), ea.double())
    for prec in ("fp32", "tf32"):
        with precision(prec) as (tol_out, _), torch.no_grad():
            xg, Zg = m(node_feat=data["node_feat"].to(DEV), node_loc=x.to(DEV), node_vel=data["vel_0"].to(DEV), edge_index=g,
                       data_batch=data["batch"].to(DEV), loc_mean=data["loc_mean"].to(DEV), edge_attr=None)
        # the update x' - x is what the layers compute; coordinates themselves are O(10) Angstrom
        from tests.gpu_util import update_err
        assert update_err(xg.cpu(), x64, x) < tol_out, (prec, update_err(xg.cpu(), x64, x))
        assert update_err(Zg.cpu(), Z64, data["loc_mean"]) < tol_out, prec
