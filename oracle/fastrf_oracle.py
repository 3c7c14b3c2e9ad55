"""CPU restatement of the reference's FastRF (models/FastRF.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.  Pinned to golden vectors
generated from the UNMODIFIED reference (oracle/make_golden.py, tests/golden/rf_*.npz; checked by
tests/test_oracle_golden.py).  Same conventions as oracle/fastegnn_oracle.py, whose primitives it reuses; every
step cites the models/FastRF.py line it follows.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor

from .fastegnn_oracle import (OracleConfig, _coord_head, _linear_init, _mlp2, global_mean_pool, segment_mean_rows)


def make_params(cfg: OracleConfig, seed: int, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Parameters drawn as the reference FastRF constructor draws them under torch.manual_seed(seed)
    (models/FastRF.py:219-225 for the stack, :27-89 for a layer: edge_mlp, edge_mlp_virtual, [att], three coordinate
    heads (1-wide xavier layer first, :56-60), coord_mlp_vel = Linear(1,H)/Linear(H,1), [gravity_mlp])."""
    torch.manual_seed(seed)
    H, C = cfg.hidden_nf, cfg.virtual_channels
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    sd["virtual_node_feat"] = torch.randn(1, H, C)                       # :219
    w, b = _linear_init(H, cfg.node_feat_nf)                              # :220
    sd["embedding_in.weight"], sd["embedding_in.bias"] = w, b

    def two_layer(prefix, in_f, out_last):
        w0, b0 = _linear_init(H, in_f)
        sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"] = w0, b0
        w2, b2 = _linear_init(out_last, H)
        sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"] = w2, b2

    def coord_head(prefix):
        last = torch.nn.Linear(H, 1, bias=False)
        torch.nn.init.xavier_uniform_(last.weight, gain=0.001)
        w0, b0 = _linear_init(H, H)
        sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"] = w0, b0
        sd[f"{prefix}.2.weight"] = last.weight.detach().clone()

    for l in range(cfg.n_layers):
        p = f"gcl_{l}"
        two_layer(f"{p}.edge_mlp", 2 * H + 1 + cfg.edge_attr_nf, H)       # :27-32
        two_layer(f"{p}.edge_mlp_virtual", 2 * H + 1 + C, H)              # :34-39
        if cfg.attention:                                                 # :42-51
            w, b = _linear_init(1, H)
            sd[f"{p}.att_mlp.0.weight"], sd[f"{p}.att_mlp.0.bias"] = w, b
            w, b = _linear_init(1, H)
            sd[f"{p}.att_mlp_virtual.0.weight"], sd[f"{p}.att_mlp_virtual.0.bias"] = w, b
        coord_head(f"{p}.coord_mlp_r")                                    # :69
        coord_head(f"{p}.coord_mlp_r_virtual")                            # :70
        coord_head(f"{p}.coord_mlp_v_virtual")                            # :71
        two_layer(f"{p}.coord_mlp_vel", 1, 1)                             # :74-78  Linear(1,H), Linear(H,1)
        if cfg.gravity is not None:                                       # :82-88
            two_layer(f"{p}.gravity_mlp", H, 1)
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


def layer_forward(params: Dict[str, Tensor], p: str, cfg: OracleConfig, h: Tensor, edge_index: Tensor, x: Tensor,
                  v: Tensor, Z: Tensor, S: Tensor, batch: Tensor, edge_attr: Tensor):
    """One FastRF ``E_GCL_vel.forward`` (models/FastRF.py:152-186): returns (x', Z'); h and S pass through (:186)."""
    N, H = h.shape
    C = cfg.virtual_channels
    vnorm = torch.norm(v, p=2, dim=-1).unsqueeze(-1).detach()            # :165
    row, col = edge_index[0], edge_index[1]
    d = x[row] - x[col]                                                   # :141-149
    q = (d * d).sum(1, keepdim=True)
    if cfg.normalize:
        d = d / (q.sqrt().detach() + cfg.eps)
    D = Z[batch] - x.unsqueeze(-1)                                        # :169
    rho = torch.norm(D, p=2, dim=1, keepdim=True)                         # :170
    m = _mlp2(params, f"{p}.edge_mlp", torch.cat([h[row], h[col], q, edge_attr], dim=1), True)      # :173, :91-97
    if cfg.attention:
        m = m * torch.sigmoid(F.linear(m, params[f"{p}.att_mlp.0.weight"], params[f"{p}.att_mlp.0.bias"]))
    xbar = global_mean_pool(x, batch, Z.size(0))                          # :175
    Zc = Z - xbar.unsqueeze(-1)                                           # :176
    M = torch.einsum("bac,bad->bcd", Zc, Zc)                              # :177
    feat = torch.cat([h.unsqueeze(-1).expand(-1, -1, C), S[batch], rho, M[batch]], dim=1)           # :100-104
    u = _mlp2(params, f"{p}.edge_mlp_virtual", feat.permute(0, 2, 1), True)
    if cfg.attention:
        u = u * torch.sigmoid(F.linear(u, params[f"{p}.att_mlp_virtual.0.weight"],
                                       params[f"{p}.att_mlp_virtual.0.bias"]))
    u = u.permute(0, 2, 1)
    # coord_model_vel, :112-140
    x_new = x + segment_mean_rows(d * _coord_head(params, f"{p}.coord_mlp_r", m, cfg.tanh), row, N)
    s_xv = _coord_head(params, f"{p}.coord_mlp_r_virtual", u.permute(0, 2, 1), cfg.tanh).permute(0, 2, 1)
    x_new = x_new + torch.mean(-D * s_xv, dim=-1)
    x_new = x_new + v * _mlp2(params, f"{p}.coord_mlp_vel", vnorm, False)                            # :135
    if cfg.gravity is not None:
        g = torch.as_tensor(cfg.gravity, dtype=x.dtype, device=x.device)
        x_new = x_new + _mlp2(params, f"{p}.gravity_mlp", h, False) * g                              # :138-139
    s_X = _coord_head(params, f"{p}.coord_mlp_v_virtual", u.permute(0, 2, 1), cfg.tanh).permute(0, 2, 1)
    Z_new = Z + global_mean_pool((D * s_X).reshape(N, -1), batch, Z.size(0)).reshape(-1, 3, C)      # :142-146
    return x_new, Z_new


def fastrf_forward(params: Dict[str, Tensor], cfg: OracleConfig, node_feat: Tensor, node_loc: Tensor, node_vel: Tensor,
                   edge_index: Tensor, data_batch: Tensor, loc_mean: Tensor, edge_attr: Tensor):
    """``FastRF.forward`` (models/FastRF.py:228-240): returns (x [N,3], Z [B,3,C])."""
    B = int(data_batch[-1]) + 1                                           # :230
    S = params["virtual_node_feat"].repeat(B, 1, 1)                       # :231
    Z = loc_mean                                                          # :232
    h = F.linear(node_feat, params["embedding_in.weight"], params["embedding_in.bias"])             # :234
    x = node_loc
    for l in range(cfg.n_layers):                                         # :235-239
        x, Z = layer_forward(params, f"gcl_{l}", cfg, h, edge_index, x, node_vel, Z, S, data_batch, edge_attr)
    return x, Z
