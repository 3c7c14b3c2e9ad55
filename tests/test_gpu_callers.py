"""The reference's OWN callers executed against the drop-in module: the body of equivariant_test.py (:11-62) and
utils/train.py::train_single_epoch (:23-179), extracted verbatim by oracle/extract_callers.py into oracle/_ref/callers.py
(git-ignored, built in the container from /root/reference, travels to the GPU box -- nothing here reads /root/reference).

Stand-ins (test infrastructure): a minimal torch_geometric-style ``Data`` / loader (collate = concatenate along dim 0,
offset ``edge_index`` by the cumulative node count, emit ``batch`` and ``ptr``; SURVEY.md 8(c) last row), dataset classes
named ``Simulation`` / ``NBodyDataset`` (the loop picks its MMD branch from the dataset's class NAME, :118), and a CPU
``nn.Module`` literally named ``FastEGNN`` (the loop dispatches on the model's class name, :51) around the oracle, so that
the SAME extracted loop produces the expected losses on the CPU."""
import importlib.util
import os
import random

import numpy as np
import pytest
import torch
from torch import nn

from oracle import fastegnn_oracle as orc
from tests.gpu_util import build_gpu_model, precision

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALLERS = os.path.join(ROOT, "oracle", "_ref", "callers.py")


def _callers():
    if not os.path.exists(CALLERS):
        pytest.skip("oracle/_ref/callers.py is not built: run `python oracle/extract_callers.py` (or __graft_entry__.build()) "
                    "in the container that has /root/reference")
    spec = importlib.util.spec_from_file_location("ref_callers", CALLERS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------- equivariant_test.py
@pytest.mark.parametrize("prec", ["tf32", "fp32", "tf32x3"])
def test_reference_equivariance_script_runs_on_the_drop_in(prec):
    """equivariant_test.py:11-62 with device='cpu' -> 'cuda:0' as the only textual change.  The script builds its tensors
    with torch.tensor(...) (CPU by default) -- torch.set_default_device('cuda:0') puts them where the model is, exactly
    what a user does to run the script on a GPU; utils/rotate.py's matrix (torch.from_numpy) is moved there as well.  The
    script's own `assert torch.allclose(..., atol=1e-4)` is the check; 6 seeds."""
    import fastegnn_b200
    ref = _callers()
    dev = "cuda:0"

    def random_rotate():
        return ref.random_rotate().to(dev)
    code = compile(ref.EQUIVARIANT_TEST, "equivariant_test.py[11:62]", "exec")
    with precision(prec):
        torch.set_default_device(dev)
        try:
            for seed in range(6):
                random.seed(seed)
                np.random.seed(seed)
                torch.manual_seed(seed)
                ns = dict(FastEGNN=fastegnn_b200.FastEGNN, nn=nn, torch=torch, random=random, np=np,
                          random_rotate=random_rotate, DEVICE=dev)
                exec(code, ns)                                   # raises AssertionError if equivariance fails
                assert ns["result_after_rotate"].is_cuda and ns["result_after_rotate"].shape == (10, 3)
                assert type(ns["model"]).__name__ == "FastEGNN"
        finally:
            torch.set_default_device("cpu")


# ------------------------------------------------------------------------------------- utils/train.py::train_single_epoch
class Data(dict):
    """Minimal stand-in for torch_geometric.data.Batch as the loop uses it: .to(device), .detach(), data['key']."""

    def to(self, device):
        return Data({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})

    def detach(self):
        return Data({k: (v.detach() if torch.is_tensor(v) else v) for k, v in self.items()})


def collate(graphs):
    off, ei, batch, ptr = 0, [], [], [0]
    for b, g in enumerate(graphs):
        n = g["loc_0"].size(0)
        ei.append(g["edge_index"] + off)
        batch.append(torch.full((n,), b, dtype=torch.long))
        off += n
        ptr.append(off)
    cat = lambda k: torch.cat([g[k] for g in graphs], dim=0)
    return Data(edge_index=torch.cat(ei, dim=1), edge_attr=cat("edge_attr"), loc_0=cat("loc_0"), vel_0=cat("vel_0"),
                loc_t=cat("loc_t"), node_feat=cat("node_feat"), node_attr=cat("node_attr"), loc_mean=cat("loc_mean"),
                batch=torch.cat(batch), ptr=torch.tensor(ptr))


class Simulation(list):          # datasets/simulation/dataset.py's class name: per-graph MMD sampling (utils/train.py:118-142)
    pass


class NBodyDataset(list):        # any other name: equal-sized graphs, one shared sample (:144-161)
    pass


class Loader:
    def __init__(self, dataset, batch_size):
        self.dataset, self.batch_size = dataset, batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield collate(self.dataset[i:i + self.batch_size])


def _graphs(sizes, C, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in sizes:
        x = torch.randn(n, 3, generator=g) * 1.2
        v = torch.randn(n, 3, generator=g) * 0.3
        d = torch.cdist(x, x) + torch.eye(n) * 1e9
        ei = (d < 1.6).nonzero().t().contiguous()                  # both directions, no self loops
        out.append(dict(loc_0=x, vel_0=v, loc_t=x + v + 0.05 * torch.randn(n, 3, generator=g),
                        node_feat=torch.stack([v.norm(dim=1), torch.rand(n, generator=g)], 1),
                        node_attr=torch.rand(n, 1, generator=g), edge_index=ei,
                        edge_attr=torch.rand(ei.size(1), 1, generator=g),
                        loc_mean=x.mean(0).reshape(1, 3, 1).repeat(1, 1, C)))
    return out


def _oracle_module(cfg, params):
    class FastEGNN(nn.Module):   # the loop dispatches on this NAME (utils/train.py:51,:111)
        def __init__(self):
            super().__init__()
            self.names = list(params)
            self.p = nn.ParameterList([nn.Parameter(params[k].clone()) for k in self.names])

        def forward(self, node_feat, node_loc, node_vel, edge_index, data_batch, loc_mean, edge_attr=None, node_attr=None):
            P = dict(zip(self.names, self.p))
            return orc.fastegnn_forward(P, cfg, node_feat, node_loc, node_vel, edge_index, data_batch, loc_mean, edge_attr)
    return FastEGNN()


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-7), ("tf32", 1e-5)])      # observed 4e-9 / 3.7e-6 (gpurun_out)
@pytest.mark.parametrize("dataset_cls,sizes", [(Simulation, [40, 55, 32, 47, 60, 38]), (NBodyDataset, [36] * 6)])
def test_reference_train_single_epoch_runs_unchanged_on_the_drop_in(dataset_cls, sizes, prec, tol):
    """utils/train.py::train_single_epoch, verbatim, drives the drop-in FastEGNN on cuda:0 with torch.optim.Adam and
    nn.MSELoss as main_*.py build them: two training epochs and one evaluation epoch return the same average losses as the
    same extracted loop driving the oracle on the CPU (same seeds for the loop's torch.randperm MMD samples)."""
    ref = _callers()
    C = 3
    cfg = orc.OracleConfig(node_feat_nf=2, edge_attr_nf=2, hidden_nf=64, virtual_channels=C, n_layers=4,
                           gravity=[0, -1, 0] if dataset_cls is Simulation else None)
    params = orc.make_params(cfg, 77)
    orc.rescale_coord_heads(params, 100.0)
    data = dataset_cls(_graphs(sizes, C, seed=3))
    loader = Loader(data, batch_size=3)
    dev = "cuda:0"
    with precision(prec):
        gpu_model = build_gpu_model(cfg, params, dev)
        cpu_model = _oracle_module(cfg, params)
        assert type(gpu_model).__name__ == type(cpu_model).__name__ == "FastEGNN"
        opt_g = torch.optim.Adam(gpu_model.parameters(), lr=5e-4, weight_decay=1e-12)      # main_nbody.py:137
        opt_c = torch.optim.Adam(cpu_model.parameters(), lr=5e-4, weight_decay=1e-12)
        loss = nn.MSELoss()
        got, want = [], []
        for epoch, backprop in ((1, True), (2, True), (3, False)):
            torch.manual_seed(100 + epoch)
            got.append(ref.train_single_epoch(gpu_model, loader, opt_g, loss, 1.5, 0.01, epoch, backprop,
                                              "train" if backprop else "valid", 3, device=dev))
            torch.manual_seed(100 + epoch)
            want.append(ref.train_single_epoch(cpu_model, loader, opt_c, loss, 1.5, 0.01, epoch, backprop,
                                               "train" if backprop else "valid", 3, device="cpu"))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/callers_train_single_epoch.txt", "a") as f:
        f.write(f"{dataset_cls.__name__} [{prec}]: drop-in {got}  oracle {want}\n")
    for e, (a, b) in enumerate(zip(got, want)):
        # epoch 1 compares one forward/backward per batch; later epochs also carry Adam's sign-like first updates
        bound = tol * (1 if e == 0 else 10) * max(1.0, abs(b))
        assert abs(a - b) <= bound, (e, a, b)
    assert gpu_model.training is False                              # the loop's last call was model.eval() (:24-27)
