"""The VNEGNN sibling (A2A / A2V / V2A stages, models/VNEGNN.py:28-375) on the GPU through the same phase
kernels as FastEGNN: against golden vectors of the unmodified reference file and against the fp64 oracle on a larger
seeded batch, in the fp32 and the default (TF32) arithmetic."""
import numpy as np
import pytest
import torch

from oracle import vnegnn_oracle as vno
from tests.gpu_util import precision, rel_err, update_err
from tests.test_vnegnn_cpu import CASES, load_vn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(params, C, n_layers, normalize=False, tanh=False):
    from fastegnn_b200 import VNEGNN
    m = VNEGNN(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=C, device=DEV,
               n_layers=n_layers, normalize=normalize, tanh=tanh)
    m.load_state_dict({k: v.to(DEV) for k, v in params.items()})
    return m


def _run(m, inp):
    g = {k: v.to(DEV) for k, v in inp.items()}
    x, Z = g["node_loc"].clone().requires_grad_(True), g["loc_mean"].clone().requires_grad_(True)
    xo, Zo = m(node_feat=g["node_feat"], node_loc=x, edge_index=g["edge_index"], data_batch=g["data_batch"],
               virtual_node_loc=Z, edge_attr=g["edge_attr"], node_attr=None)
    ((xo * g["wx"]).sum() + (Zo * g["wz"]).sum()).backward()
    torch.cuda.synchronize()
    return dict(x=xo.detach().cpu(), Z=Zo.detach().cpu(), gx=x.grad.cpu(), gZ=Z.grad.cpu(),
                gp={k: (None if p.grad is None else p.grad.cpu()) for k, p in m.named_parameters()})


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("name", CASES)
def test_vnegnn_golden_vectors_from_reference(name, prec):
    arr, params = load_vn(name)
    inp = {k[3:]: torch.from_numpy(v) for k, v in arr.items() if k.startswith("in_")}
    C = inp["loc_mean"].shape[2]
    with precision(prec) as tol:
        res = _run(_build(params, C, int(arr["n_layers"]), bool(arr["normalize"]), bool(arr["tanh"])), inp)
    assert update_err(res["x"], torch.from_numpy(arr["out_x"]), inp["node_loc"]) < tol.out + 4e-6
    assert update_err(res["Z"], torch.from_numpy(arr["out_Z"]), inp["loc_mean"]) < tol.out + 4e-6
    noise = 3e-2 if bool(arr["normalize"]) else 0.0       # the reference's own self-loop cancellation noise (normalize=True)
    assert rel_err(res["gx"], torch.from_numpy(arr["g_x"])) < tol.gin + 2e-5 + noise
    assert rel_err(res["gZ"], torch.from_numpy(arr["g_Z"])) < tol.gin + 2e-5 + noise
    none = sorted(k for k, g in res["gp"].items() if g is None or float(g.abs().max()) == 0.0)
    assert set(arr["grad_none"].tolist()) <= set(none)     # parameters of the discarded last h update get no (or zero) gradient
    for k, g in res["gp"].items():
        if "gp_" + k in arr:
            assert rel_err(g, torch.from_numpy(arr["gp_" + k])) < tol.gw + 2e-5 + noise, k


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_vnegnn_seeded_batch_against_oracle(prec):
    from tests.gpu_util import make_graph_case
    _, _, inp = make_graph_case(seed=51, sizes=[200, 150], deg=8, C=3, L=1)
    arr, params = load_vn("vn_c3_batch2")                  # the reference's own initialisation (2 layers, C = 3)
    params = {k: v.clone() for k, v in params.items()}
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    x64, Z64 = inp["node_loc"].double().requires_grad_(True), inp["loc_mean"].double().requires_grad_(True)
    xo, Zo = vno.vnegnn_forward(p64, 2, inp["node_feat"].double(), x64, inp["edge_index"], inp["data_batch"], Z64,
                                inp["edge_attr"].double())
    ((xo * inp["wx"].double()).sum() + (Zo * inp["wz"].double()).sum()).backward()
    with precision(prec) as tol:
        res = _run(_build(params, 3, 2), inp)
    assert update_err(res["x"], xo.detach(), inp["node_loc"]) < tol.out
    assert update_err(res["Z"], Zo.detach(), inp["loc_mean"]) < tol.out
    assert rel_err(res["gx"], x64.grad) < tol.gin and rel_err(res["gZ"], Z64.grad) < tol.gin
    for k, g in res["gp"].items():
        if p64[k].grad is not None:
            assert rel_err(g, p64[k].grad) < tol.gw, k
