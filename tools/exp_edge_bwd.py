"""Timing experiments on the fused edge backward (FEGNN_EXP bits; see EdgeArgs.exp).  Run once per FEGNN_EXP value."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fastegnn_b200 import _lib as L
from fastegnn_b200.ops import CsrGraph, SavedBlock, layer_ptrs, make_dims
from fastegnn_b200 import FastEGNN

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "water3d"
data, _ = bench.make_workload(which, 0, 0)
t = {k: v.to(dev) for k, v in data.items() if torch.is_tensor(v)}
N, E, B, Cc = t["loc_0"].size(0), t["edge_index"].size(1), data["n_graphs"], data["C"]
torch.manual_seed(0)
model = FastEGNN(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=Cc, device=dev, gravity=data["gravity"])
graph = CsrGraph(t["edge_index"], t["batch"], t["edge_attr"], B)
dims = make_dims(N, N, E, B, Cc, 2, L.F_GRAVITY if data["gravity"] else 0, data["gravity"])
named = dict(model.named_parameters())
ptrs = layer_ptrs(named, "gcl_0")
sv = SavedBlock(dims, dev)
sv.view("P", (N, 64)).normal_()
sv.view("Q", (N, 64)).normal_()
gv = {k: torch.zeros_like(p) for k, p in named.items() if k.startswith("gcl_0.")}
gr = layer_ptrs(gv, "gcl_0")
gm, gt = torch.randn(N, 64, device=dev), torch.randn(N, 3, device=dev)
gP, gQ, gx = torch.empty(N, 64, device=dev), torch.empty(N, 64, device=dev), torch.zeros(N, 3, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = []
for mode in tuple(int(m) for m in os.environ.get("FEGNN_EXP_MODES", "4,5,7,8").split(",")):
    L.set_mode("edge_backward", mode)
    run = lambda: L.check(L.lib.fegnn_edge_backward(C.byref(dims), C.byref(graph.c), C.byref(ptrs), C.byref(gr), L.ptr(t["loc_0"]),
                                                    C.byref(sv.c), L.ptr(gm), L.ptr(gt), L.ptr(gP), L.ptr(gQ), L.ptr(gx), st))
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    res.append(f"mode {mode}: {sum(ts) / len(ts) * 1e3:.1f} us")
print(f"{which} E={E} FEGNN_EXP={os.environ.get('FEGNN_EXP', '0')}: " + "  ".join(res))
