"""CPU restatement of the reference's VNEGNN sibling (models/VNEGNN.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module.  PINNED to golden vectors generated from the UNMODIFIED reference
(oracle/make_golden_vn.py -> tests/golden/vn_*.npz; checked by tests/test_vnegnn_cpu.py).  Functional form over the
reference's state_dict; every step cites the models/VNEGNN.py line it follows.  Reuses the primitives of
oracle/fastegnn_oracle.py (global_mean_pool stand-in, segment sums)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor

from .fastegnn_oracle import _coord_head, _mlp2, global_mean_pool, segment_mean_rows, segment_sum_rows


def a2a(p: Dict[str, Tensor], pre: str, h, edge_index, x, edge_attr, normalize=False, tanh=False, eps=1e-8):
    """EGCL_A2A.forward (:126-134)."""
    row, col = edge_index[0], edge_index[1]
    d = x[row] - x[col]                                                   # coord2radial, :114-123
    q = (d * d).sum(1, keepdim=True)
    if normalize:
        d = d / (q.sqrt().detach() + eps)
    feats = [h[row], h[col], q] + ([edge_attr] if edge_attr is not None else [])
    m = _mlp2(p, f"{pre}.edge_mlp", torch.cat(feats, dim=1), True)        # edge_model, :72-81
    x_new = x + segment_mean_rows(d * _coord_head(p, f"{pre}.coord_mlp", m, tanh), row, x.size(0))   # coord_model, :99-109
    agg = segment_sum_rows(m, row, h.size(0))                             # node_model, :84-96 -- SUM of the messages
    h_new = h + _mlp2(p, f"{pre}.node_mlp", torch.cat([h, agg], dim=1), False)
    return h_new, x_new


def _virtual_messages(p, pre, h, x, S, Z, batch):
    D = Z[batch] - x.unsqueeze(-1)                                        # [N,3,C]   :211 / :319
    rho = torch.norm(D, p=2, dim=1, keepdim=True)                         # [N,1,C]   :212 / :320
    C = S.size(2)
    feat = torch.cat([h.unsqueeze(-1).repeat(1, 1, C), S[batch], rho], dim=1)      # :184-186 / :286-288
    u = _mlp2(p, f"{pre}.edge_mlp", feat.permute(0, 2, 1), True).permute(0, 2, 1)  # [N,H,C]
    return D, u


def a2v(p, pre, h, x, S, Z, batch, tanh=False):
    """EGCL_A2V.forward (:204-225): S [B,H,C], Z [B,3,C]."""
    B, Hh, C = S.shape
    D, u = _virtual_messages(p, pre, h, x, S, Z, batch)
    trans = D * _coord_head(p, f"{pre}.coord_mlp", u.permute(0, 2, 1), tanh).permute(0, 2, 1)       # :195
    Z_new = Z + global_mean_pool(trans.reshape(trans.size(0), -1), batch, B).reshape(-1, 3, C)      # :196-197
    agg = global_mean_pool(u.reshape(u.size(0), -1), batch, B).reshape(-1, Hh, C)                   # :203-204
    out = _mlp2(p, f"{pre}.node_mlp", torch.cat([S, agg], dim=1).permute(0, 2, 1), False).permute(0, 2, 1)
    return S + out, Z_new                                                                            # :208-209


def v2a(p, pre, S, Z, h, x, batch, tanh=False):
    """EGCL_V2A.forward (:318-328)."""
    D, u = _virtual_messages(p, pre, h, x, S, Z, batch)
    s = _coord_head(p, f"{pre}.coord_mlp", u.permute(0, 2, 1), tanh).permute(0, 2, 1)
    x_new = x + torch.mean(-D * s, dim=-1)                                # coord_model_V2A, :297-301
    h_new = h + _mlp2(p, f"{pre}.node_mlp", torch.cat([h, torch.mean(u, dim=-1)], dim=1), False)    # node_model_V2A, :304-315
    return h_new, x_new


def vnegnn_forward(p: Dict[str, Tensor], n_layers: int, node_feat, node_loc, edge_index, data_batch, virtual_node_loc,
                   edge_attr=None, normalize=False, tanh=False):
    """VNEGNN.forward (:360-375): returns (node_loc [N,3], virtual_node_loc [B,3,C])."""
    B = int(data_batch[-1]) + 1                                           # :362
    S = p["virtual_node_feat"].repeat(B, 1, 1)                            # :363
    Z = virtual_node_loc
    h = F.linear(node_feat, p["embedding_in.weight"], p["embedding_in.bias"])   # :366
    x = node_loc
    for i in range(n_layers):                                             # :368-371
        h, x = a2a(p, f"A2A_{i}", h, edge_index, x, edge_attr, normalize, tanh)
        S, Z = a2v(p, f"A2V_{i}", h, x, S, Z, data_batch, tanh)
        h, x = v2a(p, f"V2A_{i}", S, Z, h, x, data_batch, tanh)
    return x, Z
