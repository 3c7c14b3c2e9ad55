"""CPU oracle for the FastEGNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, on plain CPU torch ops, the algorithm of the reference's
``models/FastEGNN.py`` (layer ``E_GCL_vel`` + stack ``FastEGNN``) and of the MMD
regulariser in ``utils/train.py``.  It exists so that the CUDA path in
``fastegnn_b200/`` can be checked for parity on machines where ``/root/reference``
is absent (the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package never does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the *unmodified*
reference module from ``/root/reference`` (with a stand-in for the one
third-party function it needs, ``torch_geometric.nn.global_mean_pool``) and
stores seeded input/output/gradient vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against them.  The reference
itself carries no golden vectors (``equivariant_test.py`` is unseeded), and the
``global_mean_pool`` semantics (PyG 2.5.2: scatter-sum / clamped count) is
restated from its published behaviour -- no reference test pins that function
numerically.

The restatement is functional (a flat ``dict`` of tensors keyed by the
reference's ``state_dict`` names) and keeps the reference's op chain (gather,
concat, Linear, scatter_add with an expanded index, including the ``ones_like``
count scatter of ``unsorted_segment_mean``) so that timing it on host cores is a
fair stand-in for the reference's own CPU path.

All ``file:line`` citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config
@dataclass
class OracleConfig:
    """Constructor arguments of the reference model (models/FastEGNN.py:227-228)."""
    node_feat_nf: int = 2
    node_attr_nf: int = 0
    edge_attr_nf: int = 2
    hidden_nf: int = 64
    virtual_channels: int = 3
    n_layers: int = 4
    residual: bool = True
    attention: bool = False
    normalize: bool = False
    tanh: bool = False
    gravity: Optional[Sequence[float]] = None
    eps: float = 1e-8                      # models/FastEGNN.py:21
    coords_agg: str = "mean"               # E_GCL_vel kwarg (:12,:124-131); FastEGNN never passes it (:261)


# --------------------------------------------------------------------------- parameters
def _linear_init(out_f: int, in_f: int, bias: bool = True) -> List[Tensor]:
    """torch.nn.Linear's default initialisation, drawn from the global RNG in the
    same order (weight first, then bias)."""
    lin = torch.nn.Linear(in_f, out_f, bias=bias)
    out = [lin.weight.detach().clone()]
    if bias:
        out.append(lin.bias.detach().clone())
    return out


def make_params(cfg: OracleConfig, seed: int, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Draw parameters exactly as the reference constructor does under
    ``torch.manual_seed(seed)``: same tensors, same order of RNG consumption
    (models/FastEGNN.py:256-262 for the stack, :28-99 for one layer; note the
    coordinate heads create their 1-wide xavier layer *before* the HxH layer,
    :56-60).  Keys and shapes equal the reference ``state_dict``."""
    torch.manual_seed(seed)
    H, C = cfg.hidden_nf, cfg.virtual_channels
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    sd["virtual_node_feat"] = torch.randn(1, H, C)                       # :256
    w, b = _linear_init(H, cfg.node_feat_nf)                              # :257
    sd["embedding_in.weight"], sd["embedding_in.bias"] = w, b

    def two_layer(prefix, in_f, out_last, last_bias=True):
        w0, b0 = _linear_init(H, in_f)
        sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"] = w0, b0
        t = _linear_init(out_last, H, bias=last_bias)
        sd[f"{prefix}.2.weight"] = t[0]
        if last_bias:
            sd[f"{prefix}.2.bias"] = t[1]

    def coord_head(prefix):
        last = torch.nn.Linear(H, 1, bias=False)                          # :56
        torch.nn.init.xavier_uniform_(last.weight, gain=0.001)            # :57
        w0, b0 = _linear_init(H, H)                                       # :60
        sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"] = w0, b0
        sd[f"{prefix}.2.weight"] = last.weight.detach().clone()

    for l in range(cfg.n_layers):
        p = f"gcl_{l}"
        two_layer(f"{p}.edge_mlp", 2 * H + 1 + cfg.edge_attr_nf, H)       # :28-33
        two_layer(f"{p}.edge_mlp_virtual", 2 * H + 1 + C, H)              # :35-40
        if cfg.attention:                                                 # :43-52
            w, b = _linear_init(1, H)
            sd[f"{p}.att_mlp.0.weight"], sd[f"{p}.att_mlp.0.bias"] = w, b
            w, b = _linear_init(1, H)
            sd[f"{p}.att_mlp_virtual.0.weight"], sd[f"{p}.att_mlp_virtual.0.bias"] = w, b
        coord_head(f"{p}.coord_mlp_r")                                    # :69
        coord_head(f"{p}.coord_mlp_r_virtual")                            # :70
        coord_head(f"{p}.coord_mlp_v_virtual")                            # :71
        two_layer(f"{p}.coord_mlp_vel", H, 1)                             # :74-78
        if cfg.gravity is not None:                                       # :82-87
            two_layer(f"{p}.gravity_mlp", H, 1)
        two_layer(f"{p}.node_mlp", H + H + C * H + cfg.node_attr_nf, H)   # :89-93
        two_layer(f"{p}.node_mlp_virtual", 2 * H, H)                      # :95-99
    # the reference's module order in state_dict differs from creation order
    # (coord heads register after att); dict order is irrelevant for lookups.
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


def rescale_coord_heads(params: Dict[str, Tensor], gain: float = 1000.0) -> None:
    """In place: multiply the xavier(gain=0.001) coordinate-head output layers so
    that the x-path has natural magnitude (SURVEY.md section 4 caveat)."""
    for k in params:
        if k.endswith(".2.weight") and ("coord_mlp_r" in k or "coord_mlp_v_virtual" in k):
            params[k].mul_(gain)


# --------------------------------------------------------------------------- primitives
def global_mean_pool(x: Tensor, batch: Tensor, size: Optional[int] = None) -> Tensor:
    """PyG 2.5.2 ``global_mean_pool`` for 2-D ``x`` (third-party; call sites
    models/FastEGNN.py:148,170,212): scatter-sum over ``batch`` divided by the
    per-graph node count clamped to >= 1."""
    B = int(batch.max()) + 1 if size is None else size
    out = x.new_zeros(B, x.size(1))
    out.scatter_add_(0, batch.unsqueeze(-1).expand(-1, x.size(1)), x)
    cnt = x.new_zeros(B).scatter_add_(0, batch, x.new_ones(x.size(0))).clamp(min=1)
    return out / cnt.unsqueeze(-1)


def segment_mean_rows(data: Tensor, seg: Tensor, n_seg: int) -> Tensor:
    """models/FastEGNN.py:287-294 (``unsorted_segment_mean``): sum of ``data`` rows
    per segment divided by a count obtained by scattering ``ones_like(data)``,
    clamped to >= 1, so empty segments yield 0."""
    idx = seg.unsqueeze(-1).expand(-1, data.size(1))
    tot = data.new_zeros(n_seg, data.size(1)).scatter_add_(0, idx, data)
    cnt = data.new_zeros(n_seg, data.size(1)).scatter_add_(0, idx, torch.ones_like(data))
    return tot / cnt.clamp(min=1)


def segment_sum_rows(data: Tensor, seg: Tensor, n_seg: int) -> Tensor:
    """models/FastEGNN.py:279-284 (``unsorted_segment_sum``)."""
    idx = seg.unsqueeze(-1).expand(-1, data.size(1))
    return data.new_zeros(n_seg, data.size(1)).scatter_add_(0, idx, data)


def _mlp2(params, prefix, x, final_act: bool):
    """Linear -> SiLU -> Linear [-> SiLU] (models/FastEGNN.py:28-33 et al.)."""
    y = F.silu(F.linear(x, params[f"{prefix}.0.weight"], params[f"{prefix}.0.bias"]))
    y = F.linear(y, params[f"{prefix}.2.weight"], params.get(f"{prefix}.2.bias"))
    return F.silu(y) if final_act else y


def _coord_head(params, prefix, x, use_tanh: bool):
    """Linear -> SiLU -> Linear(H,1,no bias) [-> tanh] (models/FastEGNN.py:55-66)."""
    y = _mlp2(params, prefix, x, final_act=False)
    return torch.tanh(y) if use_tanh else y


# --------------------------------------------------------------------------- one layer
def layer_forward(params: Dict[str, Tensor], p: str, cfg: OracleConfig,
                  h: Tensor, edge_index: Tensor, x: Tensor, v: Tensor,
                  Z: Tensor, S: Tensor, batch: Tensor,
                  edge_attr: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """One ``E_GCL_vel.forward`` (models/FastEGNN.py:192-223).

    h [N,H], x,v [N,3], Z [B,3,C], S [B,H,C]; returns (h', x', S', Z')."""
    N, H = h.shape
    C = cfg.virtual_channels
    row, col = edge_index[0], edge_index[1]

    # coord2radial, :180-189 -- SQUARED distance
    d = x[row] - x[col]
    q = (d * d).sum(1, keepdim=True)
    if cfg.normalize:
        d = d / (q.sqrt().detach() + cfg.eps)

    # virtual geometry, :206-207 -- plain (not squared) norm
    D = Z[batch] - x.unsqueeze(-1)                                        # [N,3,C]
    rho = torch.norm(D, p=2, dim=1, keepdim=True)                         # [N,1,C]

    # edge_model, :102-108
    m = _mlp2(params, f"{p}.edge_mlp", torch.cat([h[row], h[col], q, edge_attr], dim=1), True)
    if cfg.attention:
        m = m * torch.sigmoid(F.linear(m, params[f"{p}.att_mlp.0.weight"], params[f"{p}.att_mlp.0.bias"]))

    # centred virtual Gram matrix, :212-214
    xbar = global_mean_pool(x, batch, Z.size(0))                          # [B,3]
    Zc = Z - xbar.unsqueeze(-1)
    M = torch.einsum("bac,bad->bcd", Zc, Zc)                              # [B,C,C]

    # edge_mode_virtual, :111-119 -- per (node, channel) input [h_i ; S_bc ; rho_ic ; M_b[:,c]]
    feat = torch.cat([h.unsqueeze(-1).expand(-1, -1, C), S[batch], rho, M[batch]], dim=1)   # [N,2H+1+C,C]
    u = _mlp2(params, f"{p}.edge_mlp_virtual", feat.permute(0, 2, 1), True)                # [N,C,H]
    if cfg.attention:
        u = u * torch.sigmoid(F.linear(u, params[f"{p}.att_mlp_virtual.0.weight"],
                                       params[f"{p}.att_mlp_virtual.0.bias"]))
    u = u.permute(0, 2, 1)                                                                 # [N,H,C]

    # coord_model_vel, :122-144 (FastEGNN never passes coords_agg, :261: 'mean'; the layer class accepts 'sum', :124-131)
    trans = d * _coord_head(params, f"{p}.coord_mlp_r", m, cfg.tanh)
    if cfg.coords_agg == "sum":
        x_new = x + segment_sum_rows(trans, row, N)
    elif cfg.coords_agg == "mean":
        x_new = x + segment_mean_rows(trans, row, N)
    else:
        raise Exception('Wrong coords_agg parameter')                                      # :131
    s_xv = _coord_head(params, f"{p}.coord_mlp_r_virtual", u.permute(0, 2, 1), cfg.tanh).permute(0, 2, 1)
    x_new = x_new + torch.mean(-D * s_xv, dim=-1)
    x_new = x_new + _mlp2(params, f"{p}.coord_mlp_vel", h, False) * v
    if cfg.gravity is not None:
        g = torch.as_tensor(cfg.gravity, dtype=x.dtype, device=x.device)                   # :259 (int64 there; promoted)
        x_new = x_new + _mlp2(params, f"{p}.gravity_mlp", h, False) * g

    # coord_model_virtual, :146-150
    s_X = _coord_head(params, f"{p}.coord_mlp_v_virtual", u.permute(0, 2, 1), cfg.tanh).permute(0, 2, 1)
    Z_new = Z + global_mean_pool((D * s_X).reshape(N, -1), batch, Z.size(0)).reshape(-1, 3, C)

    # node_model, :153-166 -- flatten order of u is k*C + c
    agg = torch.cat([h, segment_mean_rows(m, row, N), u.reshape(N, -1)], dim=1)
    out = _mlp2(params, f"{p}.node_mlp", agg, False)
    h_new = h + out if cfg.residual else out

    # node_model_virtual, :168-177
    pooled = global_mean_pool(u.reshape(N, -1), batch, Z.size(0)).reshape(-1, H, C)
    out = _mlp2(params, f"{p}.node_mlp_virtual", torch.cat([S, pooled], dim=1).permute(0, 2, 1), False)
    out = out.permute(0, 2, 1)
    S_new = S + out if cfg.residual else out
    return h_new, x_new, S_new, Z_new


# --------------------------------------------------------------------------- the stack
def fastegnn_forward(params: Dict[str, Tensor], cfg: OracleConfig, node_feat: Tensor,
                     node_loc: Tensor, node_vel: Tensor, edge_index: Tensor,
                     data_batch: Tensor, loc_mean: Tensor, edge_attr: Tensor,
                     return_all: bool = False):
    """``FastEGNN.forward`` (models/FastEGNN.py:265-276): returns (x [N,3], Z [B,3,C])."""
    B = int(data_batch[-1]) + 1                                           # :267
    S = params["virtual_node_feat"].repeat(B, 1, 1)                       # :268
    Z = loc_mean                                                          # :269
    h = F.linear(node_feat, params["embedding_in.weight"], params["embedding_in.bias"])   # :271
    x = node_loc
    trace = []
    for l in range(cfg.n_layers):                                         # :272-275
        h, x, S, Z = layer_forward(params, f"gcl_{l}", cfg, h, edge_index, x, node_vel, Z, S,
                                   data_batch, edge_attr)
        if return_all:
            trace.append((h, x, S, Z))
    if return_all:
        return x, Z, trace
    return x, Z


# --------------------------------------------------------------------------- MMD
def laplace_kernel(x: Tensor, y: Tensor, sigma: float) -> Tensor:
    """utils/train.py:17-20 -- exp(-||x-y||_2 / (2 sigma^2)); distance, not squared."""
    return torch.exp(-torch.cdist(x, y, p=2) / (2 * sigma * sigma))


def mmd_loss(node_loc: Tensor, virtual_loc: Tensor, data_batch: Tensor, sigma: float,
             sample_idx: Sequence[Tensor]) -> Tensor:
    """MMD regulariser of utils/train.py:111-165 with the random sample made
    explicit: ``sample_idx[b]`` are the *within-graph* indices the reference would
    have drawn with ``torch.randperm(n_b)[:num_sample]`` (:131, :152).

    node_loc [N,3]; virtual_loc [B,3,C] (as returned by the model; permuted to
    [B,C,3] at :113).  Returns l_vv - l_rv (:163)."""
    Zt = virtual_loc.permute(0, 2, 1)                                     # [B,C,3]
    B, C, _ = Zt.shape
    l_vv = node_loc.new_zeros(())
    l_rv = node_loc.new_zeros(())
    ns = None
    for b in range(B):
        xb = node_loc[data_batch == b]
        xs = xb[sample_idx[b]]
        ns = xs.size(0)
        l_vv = l_vv + laplace_kernel(Zt[b], Zt[b], sigma).sum()
        l_rv = l_rv + laplace_kernel(xs, Zt[b], sigma).sum()
    l_vv = l_vv / B / C / C                                               # :141 / :160
    l_rv = 2 * l_rv / B / ns / C                                          # :142 / :161
    return l_vv - l_rv


# --------------------------------------------------------------------------- integer artefacts
def csr_by_row(edge_index: np.ndarray, n_nodes: int):
    """Stable sort of edges by destination ``row`` (the end every reference
    aggregation reduces over, models/FastEGNN.py:127-129,156).  Returns
    (perm int32 [E], rowptr int32 [N+1], row_sorted int32 [E], col_sorted int32 [E],
    deg_clamped int32 [N]); the bit-exact yard-stick for the CUDA graph-prep."""
    row = np.asarray(edge_index[0], dtype=np.int64)
    col = np.asarray(edge_index[1], dtype=np.int64)
    perm = np.argsort(row, kind="stable").astype(np.int32)
    deg = np.bincount(row, minlength=n_nodes).astype(np.int64)
    rowptr = np.zeros(n_nodes + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(deg)
    return (perm, rowptr, row[perm].astype(np.int32), col[perm].astype(np.int32),
            np.maximum(deg, 1).astype(np.int32))


def graph_ptr(data_batch: np.ndarray, n_graphs: int) -> np.ndarray:
    """``ptr`` of a PyG batch from a non-decreasing ``batch`` vector (utils/train.py:36)."""
    cnt = np.bincount(np.asarray(data_batch, dtype=np.int64), minlength=n_graphs)
    out = np.zeros(n_graphs + 1, dtype=np.int32)
    out[1:] = np.cumsum(cnt)
    return out
