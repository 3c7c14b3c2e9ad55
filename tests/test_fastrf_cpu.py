"""CPU: oracle/fastrf_oracle.py against vectors produced by the unmodified reference models/FastRF.py
(oracle/make_golden_rf.py), and the host-side FastRF module's constructor / state_dict contract."""
import numpy as np
import pytest
import torch

from oracle import fastegnn_oracle as orc
from oracle import fastrf_oracle as rfo
from tests.helpers import case_config, case_inputs, load_case, sha

RF_CASES = ["rf_c3_gravity", "rf_c3_batch3", "rf_c2_flags", "rf_h16_full_grads"]
RF_H64 = [c for c in RF_CASES if "h16" not in c]
RTOL, ATOL = 2e-5, 2e-6


def rf_case_params(case, dtype=torch.float32):
    cfg = case_config(case)
    params = rfo.make_params(cfg, case["seed"])
    if case["gain"] != 1.0:
        orc.rescale_coord_heads(params, case["gain"])
    return cfg, {k: v.to(dtype) for k, v in params.items()}


def rf_oracle_run(cfg, params, inp, want_grads=True):
    p = {k: v.clone().requires_grad_(want_grads) for k, v in params.items()}
    leaf = {k: inp[k].clone().requires_grad_(want_grads) for k in ("node_loc", "loc_mean", "node_feat")}
    x, Z = rfo.fastrf_forward(p, cfg, leaf["node_feat"], leaf["node_loc"], inp["node_vel"], inp["edge_index"],
                              inp["data_batch"], leaf["loc_mean"], inp["edge_attr"])
    res = dict(x=x.detach(), Z=Z.detach())
    if want_grads:
        ((x * inp["wx"]).sum() + (Z * inp["wz"]).sum()).backward()
        res["gin"] = {k: t.grad for k, t in leaf.items()}
        res["gp"] = {k: t.grad for k, t in p.items()}
    return res


@pytest.mark.parametrize("name", RF_CASES)
def test_rf_parameter_replay_is_bit_exact(name):
    meta, _ = load_case(name)
    _, params = rf_case_params(meta["case"])
    assert set(params) == set(meta["keys"])
    for k, h in meta["param_sha256"].items():
        assert sha(params[k]) == h, k


@pytest.mark.parametrize("name", RF_CASES)
def test_rf_forward_backward_matches_reference(name):
    torch.set_num_threads(1)
    meta, arr = load_case(name)
    cfg, params = rf_case_params(meta["case"])
    inp = case_inputs(arr)
    res = rf_oracle_run(cfg, params, inp)
    np.testing.assert_allclose(res["x"].numpy(), arr["out_x"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(res["Z"].numpy(), arr["out_Z"], rtol=RTOL, atol=ATOL)
    for k, g in res["gin"].items():
        ref = arr[f"gin_{k}"]
        scale = np.abs(ref).max() + 1e-30
        np.testing.assert_allclose(g.numpy(), ref, rtol=1e-4, atol=2e-5 * scale, err_msg=k)
    assert meta["grad_none"] == []                        # every FastRF parameter receives a gradient
    for k, dig in meta["grad_digest"].items():
        g = res["gp"][k].double().flatten()
        scale = dig["l2"] + 1e-30
        assert abs(float(g.norm()) - dig["l2"]) <= 1e-4 * scale, k
        np.testing.assert_allclose(g[dig["idx"]].numpy(), np.array(dig["val"]), rtol=1e-3, atol=1e-4 * scale, err_msg=k)
        if f"gp_{k}" in arr:
            np.testing.assert_allclose(res["gp"][k].numpy(), arr[f"gp_{k}"], rtol=1e-3, atol=1e-5 * scale, err_msg=k)


@pytest.mark.parametrize("name", RF_H64)
def test_rf_host_module_replays_reference_parameter_stream(name):
    """models.FastRF.FastRF(...) under torch.manual_seed draws the reference's tensors, bit for bit, under the
    reference's state_dict keys (the module is a parameter container on the CPU; running it needs CUDA)."""
    from models.FastRF import FastRF
    meta, _ = load_case(name)
    case = meta["case"]
    torch.manual_seed(case["seed"])
    m = FastRF(node_feat_nf=case["node_feat_nf"], node_attr_nf=0, edge_attr_nf=case["edge_attr_nf"],
               hidden_nf=case["hidden_nf"], virtual_channels=case["virtual_channels"], device="cpu",
               n_layers=case["n_layers"], attention=case.get("attention", False), normalize=case.get("normalize", False),
               tanh=case.get("tanh", False), gravity=case.get("gravity"))
    sd = m.state_dict()
    assert list(sd.keys()) == meta["keys"]
    _, params = rf_case_params(dict(case, gain=1.0))
    for k in sd:
        assert torch.equal(sd[k], params[k]), k
    assert m.__class__.__name__ == "FastRF"               # utils/train.py:57,111 dispatch on the class name
    with pytest.raises(Exception):
        m(node_feat=torch.zeros(2, 2), node_loc=torch.zeros(2, 3), node_vel=torch.zeros(2, 3),
          edge_index=torch.zeros(2, 1, dtype=torch.long), data_batch=torch.zeros(2, dtype=torch.long),
          loc_mean=torch.zeros(1, 3, case["virtual_channels"]), edge_attr=torch.zeros(1, 2))      # no CPU fallback
