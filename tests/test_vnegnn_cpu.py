"""The VNEGNN sibling on the CPU: the oracle (oracle/vnegnn_oracle.py) against golden vectors of the UNMODIFIED
models/VNEGNN.py (oracle/make_golden_vn.py), and the drop-in module's constructor against the reference's parameters
(same RNG order -> bit-identical initialisation under a seed, same state_dict keys)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import vnegnn_oracle as vno
from tests.helpers import GOLDEN

CASES = ["vn_c3_batch2", "vn_c2_flags"]


def load_vn(name):
    arr = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    params = {k[2:]: torch.from_numpy(v) for k, v in arr.items() if k.startswith("p_")}
    return arr, params


@pytest.mark.parametrize("name", CASES)
def test_vnegnn_oracle_matches_reference(name):
    torch.set_num_threads(1)
    arr, params = load_vn(name)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    t = lambda k: torch.from_numpy(arr[k])
    x, Z = t("in_node_loc").clone().requires_grad_(True), t("in_loc_mean").clone().requires_grad_(True)
    xo, Zo = vno.vnegnn_forward(p, int(arr["n_layers"]), t("in_node_feat"), x, t("in_edge_index"), t("in_data_batch"), Z,
                                t("in_edge_attr"), bool(arr["normalize"]), bool(arr["tanh"]))
    ((xo * t("in_wx")).sum() + (Zo * t("in_wz")).sum()).backward()
    np.testing.assert_allclose(xo.detach().numpy(), arr["out_x"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(Zo.detach().numpy(), arr["out_Z"], rtol=2e-5, atol=2e-6)
    for got, key in ((x, "g_x"), (Z, "g_Z")):
        scale = np.abs(arr[key]).max() + 1e-30
        np.testing.assert_allclose(got.grad.numpy(), arr[key], rtol=1e-4, atol=2e-5 * scale, err_msg=key)
    none = sorted(k for k, v in p.items() if v.grad is None)
    assert none == sorted(arr["grad_none"].tolist())
    for k, v in p.items():
        if v.grad is None:
            continue
        ref = arr["gp_" + k]
        scale = np.abs(ref).max() + 1e-30
        # normalize=True: a self-loop's +g/1e-8 and -g/1e-8 cancel only to rounding in both implementations
        np.testing.assert_allclose(v.grad.numpy(), ref, rtol=1e-3, atol=(2e-3 if bool(arr["normalize"]) else 3e-5) * scale,
                                   err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_drop_in_constructor_replays_the_reference_initialisation(name):
    from fastegnn_b200 import VNEGNN
    arr, params = load_vn(name)
    C = int(arr["in_loc_mean"].shape[2])
    torch.manual_seed(int(arr["seed"]))
    m = VNEGNN(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=C, device="cpu",
               n_layers=int(arr["n_layers"]), normalize=bool(arr["normalize"]), tanh=bool(arr["tanh"]))
    sd = m.state_dict()
    assert list(sd.keys()) == list(params.keys())                   # same keys in the same order
    for k, v in sd.items():
        assert hashlib.sha256(v.numpy().tobytes()).hexdigest() == str(arr["sha_" + k]), k
    with pytest.raises(Exception):                                    # no CPU path
        m(node_feat=torch.zeros(4, 2), node_loc=torch.zeros(4, 3), edge_index=torch.zeros(2, 3, dtype=torch.long),
          data_batch=torch.zeros(4, dtype=torch.long), virtual_node_loc=torch.zeros(1, 3, C), edge_attr=torch.zeros(3, 2))
