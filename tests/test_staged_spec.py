"""oracle/staged.py (the kernel pipeline spelled out phase by phase, hand-written
backward) must equal the oracle's autograd.  float64, CPU."""
import numpy as np
import pytest
import torch

from oracle import fastegnn_oracle as orc
from oracle import staged
from tests.helpers import H64_CASES, MODEL_CASES, case_inputs, case_params, load_case, oracle_run


@pytest.mark.parametrize("name", MODEL_CASES)
def test_staged_equals_oracle_autograd_fp64(name):
    meta, arr = load_case(name)
    case = meta["case"]
    cfg, params = case_params(case, torch.float64)
    inp = case_inputs(arr, torch.float64)
    ref = oracle_run(cfg, params, inp)

    fl = staged.Flags(cfg.attention, cfg.normalize, cfg.tanh, cfg.gravity, cfg.eps)
    sm = staged.StagedModel(params, cfg.hidden_nf, cfg.virtual_channels, cfg.edge_attr_nf, cfg.n_layers, fl)
    x, Z = sm.forward(inp["node_feat"], inp["node_loc"], inp["node_vel"], inp["edge_index"], inp["data_batch"],
                      inp["loc_mean"], inp["edge_attr"])
    assert torch.allclose(x, ref["x"], rtol=1e-10, atol=1e-11)
    assert torch.allclose(Z, ref["Z"], rtol=1e-10, atol=1e-11)
    grads, gin = sm.backward(inp["wx"], inp["wz"])
    for k in ("node_loc", "loc_mean", "node_feat"):
        s = ref["gin"][k].abs().max()
        assert torch.allclose(gin[k], ref["gin"][k], rtol=1e-8, atol=1e-10 * float(s)), k
    for k, g in ref["gp"].items():
        if g is None:
            assert k not in grads, k
            continue
        s = float(g.abs().max()) + 1e-30
        assert k in grads, k
        assert torch.allclose(grads[k], g, rtol=1e-7, atol=1e-9 * s), (k, float((grads[k] - g).abs().max()), s)
