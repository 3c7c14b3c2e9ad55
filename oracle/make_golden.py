"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python oracle/make_golden.py

What it does
  * registers a stand-in for ``torch_geometric.nn.global_mean_pool`` (the single
    third-party function ``models/FastEGNN.py:4`` needs; PyG is not installed
    here) and imports the reference module from /root/reference untouched;
  * for a list of small seeded cases builds the reference ``FastEGNN`` under
    ``torch.manual_seed(seed)``, runs forward + backward of a fixed linear
    functional of the outputs, and stores inputs, outputs, input-gradients and a
    compact digest of every parameter and parameter-gradient (sha256 of the
    parameter bytes; sum / l2 / sampled entries of each gradient; full gradients
    for the narrow H=16 case);
  * extracts the MMD block of ``utils/train.py:111-165`` with ``ast`` and
    executes it verbatim on seeded tensors (the file itself cannot be imported:
    it pulls in MDAnalysis at :7).

Test infrastructure; never imported by the product.
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle import fastegnn_oracle as orc  # noqa: E402


def _install_pyg_standin():
    def global_mean_pool(x, batch, size=None):
        B = int(batch.max()) + 1 if size is None else size
        out = x.new_zeros(B, x.size(1)).scatter_add_(0, batch.unsqueeze(-1).expand(-1, x.size(1)), x)
        cnt = x.new_zeros(B).scatter_add_(0, batch, x.new_ones(x.size(0))).clamp(min=1)
        return out / cnt.unsqueeze(-1)
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgnn.global_mean_pool = global_mean_pool
    tg.nn = tgnn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = tgnn


def _import_reference_model():
    _install_pyg_standin()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_FastEGNN", os.path.join(REF, "models", "FastEGNN.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------ seeded graphs
def make_case_inputs(case: dict):
    """Small adversarial batch: unequal graph sizes, self-loops, duplicate edges,
    at least one node with no incoming-row edges."""
    g = torch.Generator().manual_seed(case["data_seed"])
    sizes = case["graph_sizes"]
    N = sum(sizes)
    B = len(sizes)
    batch = torch.cat([torch.full((n,), b, dtype=torch.long) for b, n in enumerate(sizes)])
    x = torch.randn(N, 3, generator=g) * case.get("coord_scale", 1.5)
    v = torch.randn(N, 3, generator=g) * 0.5
    nf = torch.rand(N, case["node_feat_nf"], generator=g)
    rows, cols = [], []
    off = 0
    for n in sizes:
        e = case["edges_per_graph"]
        r = torch.randint(0, n, (e,), generator=g)
        c = torch.randint(0, n, (e,), generator=g)
        if n > 2:
            r[r == n - 1] = 0          # last node of each graph: isolated as a row
        if e >= 4:
            r[1], c[1] = r[0], c[0]    # duplicate edge
            c[2] = r[2]                # self loop
        rows.append(r + off)
        cols.append(c + off)
        off += n
    edge_index = torch.stack([torch.cat(rows), torch.cat(cols)])
    E = edge_index.size(1)
    ea = torch.rand(E, case["edge_attr_nf"], generator=g)
    C = case["virtual_channels"]
    loc_mean = torch.stack([x[batch == b].mean(0) for b in range(B)]).unsqueeze(-1).repeat(1, 1, C)
    if case.get("perturb_loc_mean", True):
        loc_mean = loc_mean + 0.3 * torch.randn(B, 3, C, generator=g)   # channels must differ to exercise M
    wx = torch.randn(N, 3, generator=g)
    wz = torch.randn(B, 3, C, generator=g)
    return dict(node_feat=nf, node_loc=x, node_vel=v, edge_index=edge_index, data_batch=batch,
                loc_mean=loc_mean, edge_attr=ea, wx=wx, wz=wz)


CASES = [
    dict(name="c3_batch3", seed=11, data_seed=101, hidden_nf=64, virtual_channels=3, n_layers=4,
         node_feat_nf=2, edge_attr_nf=2, graph_sizes=[7, 12, 5], edges_per_graph=30, gain=1000.0),
    dict(name="c3_gravity", seed=12, data_seed=102, hidden_nf=64, virtual_channels=3, n_layers=4,
         node_feat_nf=2, edge_attr_nf=2, graph_sizes=[40], edges_per_graph=300, gain=1000.0,
         gravity=[0, -1, 0]),
    dict(name="c8_two_layers", seed=13, data_seed=103, hidden_nf=64, virtual_channels=8, n_layers=2,
         node_feat_nf=2, edge_attr_nf=2, graph_sizes=[9, 9], edges_per_graph=40, gain=1000.0),
    dict(name="c1_flags", seed=14, data_seed=104, hidden_nf=64, virtual_channels=1, n_layers=2,
         node_feat_nf=2, edge_attr_nf=2, graph_sizes=[10, 6], edges_per_graph=25, gain=1000.0,
         attention=True, normalize=True, tanh=True),
    dict(name="equiv_shape_default_init", seed=15, data_seed=105, hidden_nf=64, virtual_channels=3, n_layers=4,
         node_feat_nf=1, edge_attr_nf=1, graph_sizes=[10], edges_per_graph=20, gain=1.0, coord_scale=3.0,
         perturb_loc_mean=False),
    dict(name="h16_full_grads", seed=16, data_seed=106, hidden_nf=16, virtual_channels=2, n_layers=3,
         node_feat_nf=2, edge_attr_nf=2, graph_sizes=[6, 8], edges_per_graph=20, gain=1000.0,
         gravity=[0, -1, 0]),
]


def _digest(t: torch.Tensor) -> dict:
    a = t.detach().double().flatten()
    idx = np.unique(np.linspace(0, a.numel() - 1, num=min(16, a.numel())).astype(np.int64))
    return dict(sum=float(a.sum()), l2=float(a.norm()), idx=idx.tolist(), val=a[idx].tolist())


def run_model_case(ref, case):
    torch.manual_seed(case["seed"])
    model = ref.FastEGNN(node_feat_nf=case["node_feat_nf"], node_attr_nf=0, edge_attr_nf=case["edge_attr_nf"],
                         hidden_nf=case["hidden_nf"], virtual_channels=case["virtual_channels"], device="cpu",
                         n_layers=case["n_layers"], residual=True, attention=case.get("attention", False),
                         normalize=case.get("normalize", False), tanh=case.get("tanh", False),
                         gravity=case.get("gravity"))
    # the oracle's RNG replay must give the same tensors bit for bit
    cfg = orc.OracleConfig(node_feat_nf=case["node_feat_nf"], edge_attr_nf=case["edge_attr_nf"],
                           hidden_nf=case["hidden_nf"], virtual_channels=case["virtual_channels"],
                           n_layers=case["n_layers"], attention=case.get("attention", False),
                           normalize=case.get("normalize", False), tanh=case.get("tanh", False),
                           gravity=case.get("gravity"))
    replay = orc.make_params(cfg, case["seed"])
    sd = model.state_dict()
    assert set(sd.keys()) == set(replay.keys()), (set(sd) ^ set(replay))
    for k in sd:
        assert torch.equal(sd[k], replay[k]), f"RNG replay mismatch at {k}"
    # natural-magnitude coordinate heads (SURVEY.md section 4 caveat)
    if case["gain"] != 1.0:
        with torch.no_grad():
            for k, p in model.named_parameters():
                if k.endswith(".2.weight") and ("coord_mlp_r" in k or "coord_mlp_v_virtual" in k):
                    p.mul_(case["gain"])
    inp = make_case_inputs(case)
    leaf = {k: inp[k].clone().requires_grad_(True) for k in ("node_loc", "node_vel", "loc_mean", "node_feat")}
    x, Z = model(node_feat=leaf["node_feat"], node_loc=leaf["node_loc"], node_vel=leaf["node_vel"],
                 edge_index=inp["edge_index"], data_batch=inp["data_batch"], loc_mean=leaf["loc_mean"],
                 edge_attr=inp["edge_attr"], node_attr=None)
    loss = (x * inp["wx"]).sum() + (Z * inp["wz"]).sum()
    loss.backward()

    arrays = {f"in_{k}": v.numpy() for k, v in inp.items()}
    arrays["out_x"] = x.detach().numpy()
    arrays["out_Z"] = Z.detach().numpy()
    for k, t in leaf.items():
        arrays[f"gin_{k}"] = t.grad.numpy()
    meta = dict(case=case, loss=float(loss), keys=list(sd.keys()), param_sha256={}, grad_digest={}, grad_none=[])
    for k, p in model.named_parameters():
        meta["param_sha256"][k] = hashlib.sha256(p.detach().numpy().tobytes()).hexdigest()
        if p.grad is None:
            meta["grad_none"].append(k)                 # last layer node_mlp*, 8 tensors
        else:
            meta["grad_digest"][k] = _digest(p.grad)
            if case["hidden_nf"] <= 16:
                arrays[f"gp_{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"{case['name']}.npz"), **arrays)
    with open(os.path.join(OUT, f"{case['name']}.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"{case['name']}: N={x.size(0)} E={inp['edge_index'].size(1)} loss={float(loss):.6f} "
          f"grad_none={len(meta['grad_none'])}")


# ------------------------------------------------------------------ MMD block, verbatim
def _extract_mmd_block():
    src = open(os.path.join(REF, "utils", "train.py")).read()
    tree = ast.parse(src)
    kernel_src = None
    mmd_src = None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "kernel":
            kernel_src = ast.unparse(node)
        if isinstance(node, ast.FunctionDef) and node.name == "train_single_epoch":
            for sub in ast.walk(node):
                if isinstance(sub, ast.If) and "FastRF" in ast.unparse(sub.test) and "FastEGNN" in ast.unparse(sub.test) \
                        and "loss_mmd" in ast.unparse(sub):
                    mmd_src = ast.unparse(sub)
    assert kernel_src and mmd_src
    return kernel_src, mmd_src


def run_mmd_cases():
    kernel_src, mmd_src = _extract_mmd_block()

    class FastEGNN:                       # only __class__.__name__ is inspected (utils/train.py:111)
        pass

    def dataset_named(name):
        return type(name, (), {})()

    out = {}
    meta = {}
    for tag, ds_name, sizes, C, sigma, sample, dseed in [
        ("simulation_ragged", "Simulation", [9, 14, 11], 3, 1.0, 3, 201),
        ("nbody_equal", "NBodySystemDataset", [5, 5, 5, 5], 3, 1.5, 3, 202),
        ("simulation_c8", "Simulation", [30], 8, 1.0, 3, 203),
    ]:
        g = torch.Generator().manual_seed(dseed)
        N, B = sum(sizes), len(sizes)
        loc = torch.randn(N, 3, generator=g)
        Z = torch.randn(B, 3, C, generator=g)
        batch = torch.cat([torch.full((n,), b, dtype=torch.long) for b, n in enumerate(sizes)])
        ns = {"torch": torch}
        exec(kernel_src, ns)
        ns.update(model=FastEGNN(), loc_predict=loc.clone().requires_grad_(True),
                  virtual_node_loc=Z.clone().requires_grad_(True), sample=sample, batch_size=B,
                  data={"batch": batch}, sigma=sigma, weight=1.0, loss_loc=torch.zeros(()),
                  loader=types.SimpleNamespace(dataset=dataset_named(ds_name)))
        lp, vz = ns["loc_predict"], ns["virtual_node_loc"]
        torch.manual_seed(1234)
        exec(mmd_src, ns)
        val = ns["loss_mmd"]
        val.backward()
        # replay the RNG draws to recover the sample indices the block used
        torch.manual_seed(1234)
        num_sample = min(sample * C, N)
        idx = []
        if ds_name == "Simulation":
            for b in range(B):
                idx.append(torch.randperm(sizes[b])[:num_sample])
        else:
            n = sizes[0]
            shared = torch.randperm(n)[:min(num_sample, n)]
            idx = [shared for _ in range(B)]
        out[f"{tag}_loc"] = loc.numpy()
        out[f"{tag}_Z"] = Z.numpy()
        out[f"{tag}_batch"] = batch.numpy()
        out[f"{tag}_gloc"] = lp.grad.numpy()
        out[f"{tag}_gZ"] = vz.grad.numpy()
        for b, t in enumerate(idx):
            out[f"{tag}_idx{b}"] = t.numpy()
        meta[tag] = dict(sizes=sizes, C=C, sigma=sigma, sample=sample, value=float(val))
        print(f"mmd {tag}: {float(val):.8f}")
    np.savez_compressed(os.path.join(OUT, "mmd.npz"), **out)
    with open(os.path.join(OUT, "mmd.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)      # fixed reduction order for the stored vectors
    ref = _import_reference_model()
    for case in CASES:
        run_model_case(ref, case)
    run_mmd_cases()
