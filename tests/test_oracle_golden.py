"""The oracle (oracle/fastegnn_oracle.py) against vectors produced by the unmodified
reference (oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fastegnn_oracle as orc
from tests.helpers import GOLDEN, MODEL_CASES, case_inputs, case_params, load_case, oracle_run, sha

# fp32 tolerance: oracle and reference run the same ATen ops in (nearly) the same
# order on one thread, so differences are a few ulp of the intermediate magnitudes.
RTOL, ATOL = 2e-5, 2e-6


@pytest.mark.parametrize("name", MODEL_CASES)
def test_parameter_replay_is_bit_exact(name):
    meta, _ = load_case(name)
    _, params = case_params(meta["case"])
    assert set(params) == set(meta["keys"])
    for k, h in meta["param_sha256"].items():
        assert sha(params[k]) == h, k


@pytest.mark.parametrize("name", MODEL_CASES)
def test_forward_backward_matches_reference(name):
    torch.set_num_threads(1)
    meta, arr = load_case(name)
    cfg, params = case_params(meta["case"])
    inp = case_inputs(arr)
    res = oracle_run(cfg, params, inp)
    np.testing.assert_allclose(res["x"].numpy(), arr["out_x"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(res["Z"].numpy(), arr["out_Z"], rtol=RTOL, atol=ATOL)
    for k, g in res["gin"].items():
        ref = arr[f"gin_{k}"]
        scale = np.abs(ref).max() + 1e-30
        np.testing.assert_allclose(g.numpy(), ref, rtol=1e-4, atol=2e-5 * scale, err_msg=k)
    # reference leaves the last layer's node_mlp / node_mlp_virtual without gradient
    none = sorted(k for k, g in res["gp"].items() if g is None)
    assert none == sorted(meta["grad_none"])
    assert len(none) == 8
    for k, dig in meta["grad_digest"].items():
        g = res["gp"][k].double().flatten()
        scale = dig["l2"] + 1e-30
        assert abs(float(g.norm()) - dig["l2"]) <= 1e-4 * scale, k
        np.testing.assert_allclose(g[dig["idx"]].numpy(), np.array(dig["val"]), rtol=1e-3, atol=1e-4 * scale, err_msg=k)
        if f"gp_{k}" in arr:
            np.testing.assert_allclose(res["gp"][k].numpy(), arr[f"gp_{k}"], rtol=1e-3, atol=1e-5 * scale, err_msg=k)


def test_mmd_matches_reference_block():
    meta = json.load(open(os.path.join(GOLDEN, "mmd.json")))
    arr = dict(np.load(os.path.join(GOLDEN, "mmd.npz")))
    for tag, m in meta.items():
        loc = torch.from_numpy(arr[f"{tag}_loc"]).requires_grad_(True)
        Z = torch.from_numpy(arr[f"{tag}_Z"]).requires_grad_(True)
        batch = torch.from_numpy(arr[f"{tag}_batch"])
        idx = [torch.from_numpy(arr[f"{tag}_idx{b}"]) for b in range(len(m["sizes"]))]
        val = orc.mmd_loss(loc, Z, batch, m["sigma"], idx)
        assert abs(float(val) - m["value"]) < 1e-6, tag
        val.backward()
        np.testing.assert_allclose(loc.grad.numpy(), arr[f"{tag}_gloc"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(Z.grad.numpy(), arr[f"{tag}_gZ"], rtol=1e-4, atol=1e-7)


def test_csr_by_row_is_stable_sort():
    rng = np.random.default_rng(0)
    for N, E in [(1, 0), (5, 7), (50, 400), (9, 200)]:
        ei = rng.integers(0, N, size=(2, E))
        perm, rowptr, rs, cs, deg = orc.csr_by_row(ei, N)
        t = torch.sort(torch.from_numpy(ei[0]), stable=True)
        assert np.array_equal(perm, t.indices.numpy().astype(np.int32))
        assert np.array_equal(rs, t.values.numpy().astype(np.int32))
        assert np.array_equal(cs, ei[1][perm].astype(np.int32))
        assert rowptr[0] == 0 and rowptr[-1] == E
        assert np.array_equal(np.diff(rowptr), np.bincount(ei[0], minlength=N))
        assert np.array_equal(deg, np.maximum(np.bincount(ei[0], minlength=N), 1))


def test_equivariance_property_of_oracle():
    """FastEGNN(G R + t) == FastEGNN(G) R + t  (equivariant_test.py:62, atol 1e-4)."""
    torch.manual_seed(0)
    cfg = orc.OracleConfig(node_feat_nf=1, edge_attr_nf=1, virtual_channels=3)
    params = orc.make_params(cfg, 3)
    g = torch.Generator().manual_seed(5)
    N, E = 10, 20
    x = torch.rand(N, 3, generator=g) * 10
    v = torch.rand(N, 3, generator=g) * 10
    nf = torch.rand(N, 1, generator=g) * 10
    ei = torch.randint(0, N, (2, E), generator=g)
    ea = torch.rand(E, 1, generator=g) * 10
    batch = torch.zeros(N, dtype=torch.long)
    A = torch.randn(3, 3, generator=g, dtype=torch.float64)
    R, _ = torch.linalg.qr(A)
    if torch.det(R) < 0:
        R[:, 0] = -R[:, 0]
    R = R.float()
    t = torch.randn(3, generator=g) * 5

    def run(xx, vv):
        lm = xx.mean(0).unsqueeze(-1).repeat(1, 3).unsqueeze(0)
        return orc.fastegnn_forward(params, cfg, nf, xx, vv, ei, batch, lm, ea)[0]
    a = run(x, v) @ R + t
    b = run(x @ R + t, v @ R)
    assert torch.allclose(a, b, atol=1e-4)


@pytest.mark.parametrize("name", ["layer_sum", "layer_mean_gravity"])
def test_layer_forward_matches_reference_layer(name):
    """E_GCL_vel called directly (oracle/make_golden_layer.py), which pins coords_agg='sum' (models/FastEGNN.py:124-125)."""
    from tests.helpers import load_layer_case
    torch.set_num_threads(1)
    cfg, params, arr = load_layer_case(name)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    t = lambda k: torch.from_numpy(arr[k])
    h, x, Z, S = (t(k).clone().requires_grad_(True) for k in ("in_h", "in_node_loc", "in_loc_mean", "in_S"))
    ho, xo, So, Zo = orc.layer_forward(p, "gcl_0", cfg, h, t("in_edge_index"), x, t("in_node_vel"), Z, S,
                                       t("in_data_batch"), t("in_edge_attr"))
    ((ho * t("wh")).sum() + (xo * t("in_wx")).sum() + (So * t("wS")).sum() + (Zo * t("in_wz")).sum()).backward()
    for got, key in ((ho, "out_h"), (xo, "out_x"), (So, "out_S"), (Zo, "out_Z")):
        np.testing.assert_allclose(got.detach().numpy(), arr[key], rtol=RTOL, atol=ATOL, err_msg=key)
    for got, key in ((h, "g_h"), (x, "g_x"), (S, "g_S"), (Z, "g_Z")):
        scale = np.abs(arr[key]).max() + 1e-30
        np.testing.assert_allclose(got.grad.numpy(), arr[key], rtol=1e-4, atol=2e-5 * scale, err_msg=key)
    for k, v in p.items():
        ref = arr["gp_" + k[len("gcl_0."):]]
        scale = np.abs(ref).max() + 1e-30
        np.testing.assert_allclose(v.grad.numpy(), ref, rtol=1e-3, atol=2e-5 * scale, err_msg=k)


def test_wrong_coords_agg_raises_like_the_reference():
    from tests.helpers import load_layer_case
    cfg, params, arr = load_layer_case("layer_sum")
    cfg.coords_agg = "max"
    t = lambda k: torch.from_numpy(arr[k])
    with pytest.raises(Exception, match="Wrong coords_agg parameter"):
        orc.layer_forward(params, "gcl_0", cfg, t("in_h"), t("in_edge_index"), t("in_node_loc"), t("in_node_vel"),
                          t("in_loc_mean"), t("in_S"), t("in_data_batch"), t("in_edge_attr"))


def test_segment_helpers_match_reference():
    arr = dict(np.load(os.path.join(GOLDEN, "segment_helpers.npz")))
    data, ids, n = torch.from_numpy(arr["data"]), torch.from_numpy(arr["ids"]), int(arr["num"])
    np.testing.assert_allclose(orc.segment_sum_rows(data, ids, n).numpy(), arr["sum"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.segment_mean_rows(data, ids, n).numpy(), arr["mean"], rtol=1e-6, atol=1e-6)
