import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vnegnn_oracle as vno
from tests.test_vnegnn_cpu import load_vn
from fastegnn_b200 import VNEGNN, _lib as L
from fastegnn_b200.VNEGNN import _zero_table, _mlp, H
from fastegnn_b200.layer_fn import layer_call
from fastegnn_b200.ops import CsrGraph
L.set_precision("fp32")
dev = "cuda:0"
arr, params = load_vn("vn_c3_batch2")
inp = {k[3:]: torch.from_numpy(v) for k, v in arr.items() if k.startswith("in_")}
C, B = 3, 2
m = VNEGNN(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=C, device=dev, n_layers=2)
m.load_state_dict({k: v.to(dev) for k, v in params.items()})
P = {k: v.double() for k, v in params.items()}
g = {k: v.to(dev) for k, v in inp.items()}
N = g["node_loc"].size(0)
graph = CsrGraph(g["edge_index"], g["data_batch"], g["edge_attr"], B)
no_edges = CsrGraph(g["edge_index"][:, :0].contiguous(), g["data_batch"], None, B)
zero_v = torch.zeros(N, 3, device=dev)
z1, zC = _zero_table(1, 2, dev), _zero_table(C, 0, dev)
Zd, Sd = torch.zeros(B, 3, 1, device=dev), torch.zeros(B, 1, H, device=dev)
S = m.virtual_node_feat[0].t().unsqueeze(0).expand(B, C, H).contiguous()
Z = g["loc_mean"]
h = torch.nn.functional.linear(g["node_feat"], m.embedding_in.weight, m.embedding_in.bias)
x = g["node_loc"]
# reference
Sr = P["virtual_node_feat"].repeat(B, 1, 1); Zr = inp["loc_mean"].double()
hr = torch.nn.functional.linear(inp["node_feat"].double(), P["embedding_in.weight"], P["embedding_in.bias"]); xr = inp["node_loc"].double()
err = lambda a, b: float((a.detach().cpu().double() - b).abs().max() / (b.abs().max() + 1e-30))
a2a, a2v, v2a = m.A2A_0, m.A2V_0, m.V2A_0
t = dict(z1); t.update(_mlp("edge_mlp", a2a.edge_mlp)); t.update(_mlp("coord_mlp_r", a2a.coord_mlp))
w0 = a2a.node_mlp[0].weight
t.update(_mlp("node_mlp", a2a.node_mlp, torch.cat([w0, torch.zeros(H, H, device=dev)], dim=1).contiguous()))
with torch.no_grad():
    h1, x1, _, _ = layer_call(t, L.F_NODE_SUM, None, 1, graph, h, x, zero_v, Zd, Sd)
    hr1, xr1 = vno.a2a(P, "A2A_0", hr, inp["edge_index"], xr, inp["edge_attr"].double())
    print("A2A h", err(h1, hr1), "x", err(x1, xr1))
    pad = torch.zeros(H, C, device=dev)
    t = dict(zC); t.update(_mlp("edge_mlp_virtual", a2v.edge_mlp, torch.cat([a2v.edge_mlp[0].weight, pad], 1).contiguous()))
    t.update(_mlp("coord_mlp_v_virtual", a2v.coord_mlp)); t.update(_mlp("node_mlp_virtual", a2v.node_mlp))
    hq, xq, S1, Z1 = layer_call(t, 0, None, C, no_edges, hr1.float().to(dev), xr1.float().to(dev), zero_v, Z, S)
    Sr1, Zr1 = vno.a2v(P, "A2V_0", hr1, xr1, Sr, Zr, inp["data_batch"])
    print("A2V S", err(S1.permute(0, 2, 1), Sr1), "Z", err(Z1, Zr1), "x kept", err(xq, xr1), "h kept", err(hq, hr1))
    t = dict(zC); t.update(_mlp("edge_mlp_virtual", v2a.edge_mlp, torch.cat([v2a.edge_mlp[0].weight, pad], 1).contiguous()))
    t.update(_mlp("coord_mlp_r_virtual", v2a.coord_mlp))
    w0 = v2a.node_mlp[0].weight
    w_exp = torch.cat([w0[:, :H], torch.zeros(H, H, device=dev), w0[:, H:].repeat_interleave(C, dim=1) / C], dim=1)
    t.update(_mlp("node_mlp", v2a.node_mlp, w_exp.contiguous()))
    Sin, Zin = Sr1.permute(0, 2, 1).contiguous().float().to(dev), Zr1.float().to(dev)
    h2, x2, Sq, Zq = layer_call(t, 0, None, C, no_edges, hr1.float().to(dev), xr1.float().to(dev), zero_v, Zin, Sin)
    hr2, xr2 = vno.v2a(P, "V2A_0", Sr1, Zr1, hr1, xr1, inp["data_batch"])
    print("V2A h", err(h2, hr2), "x", err(x2, xr2), "S kept", err(Sq.permute(0, 2, 1), Sr1), "Z kept", err(Zq, Zr1))
