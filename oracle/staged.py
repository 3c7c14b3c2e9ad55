"""Phase-by-phase CPU specification of the kernel pipeline -- TEST INFRASTRUCTURE.

``fastegnn_oracle.py`` restates the reference's formulation (concat + Linear,
autograd backward).  The CUDA path computes the same function through a different
but algebraically equal pipeline:

  * the first Linear of every MLP whose input is a concatenation is split into
    per-source blocks, so ``phi_e``'s 131-wide GEMM becomes two per-NODE products
    ``P = W_src h + b`` and ``Q = W_tgt h`` plus a rank-(1+Fe) per-edge update;
  * the backward is written out by hand, with per-edge quantities recomputed
    instead of stored.

This module spells that pipeline out with explicit formulas (no autograd), one
function per kernel phase, using the same intermediate names as
``fastegnn_b200/csrc``.  ``tests/test_staged_spec.py`` proves it equal to the
oracle's autograd in float64; the GPU tests then compare every CUDA phase against
the matching function here, which localises a failing kernel.

Shapes: h [N,H]; x,v [Nl,3] (Nl >= N: owned + halo rows in the partitioned path);
Z [B,3,C]; S [B,C,H] (channel-major -- the reference stores [B,H,C]); row/col [E]
with row < N sorted ascending (CSR order) and col < Nl.
Citations are to /root/reference/models/FastEGNN.py.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

T = torch.Tensor


def silu(z):
    return z * torch.sigmoid(z)


def dsilu(z):
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


class LayerWeights:
    """Views of one layer's reference-named tensors split the way the kernels use them."""

    def __init__(self, sd: Dict[str, T], p: str, H: int, C: int, Fe: int, attention: bool, gravity: bool):
        g = lambda k: sd[f"{p}.{k}"]
        W1 = g("edge_mlp.0.weight")                      # [H, 2H+1+Fe]  (:29)
        self.Ws, self.Wt = W1[:, :H], W1[:, H:2 * H]
        self.wq, self.Wa, self.b1 = W1[:, 2 * H], W1[:, 2 * H + 1:], g("edge_mlp.0.bias")
        self.W2, self.b2 = g("edge_mlp.2.weight"), g("edge_mlp.2.bias")
        self.W3, self.b3 = g("coord_mlp_r.0.weight"), g("coord_mlp_r.0.bias")
        self.w4 = g("coord_mlp_r.2.weight")[0]
        V1 = g("edge_mlp_virtual.0.weight")              # [H, 2H+1+C]   (:36)
        self.V1h, self.V1s = V1[:, :H], V1[:, H:2 * H]
        self.vr, self.V1m, self.c1 = V1[:, 2 * H], V1[:, 2 * H + 1:], g("edge_mlp_virtual.0.bias")
        self.V2, self.c2 = g("edge_mlp_virtual.2.weight"), g("edge_mlp_virtual.2.bias")
        self.Wxv, self.bxv = g("coord_mlp_r_virtual.0.weight"), g("coord_mlp_r_virtual.0.bias")
        self.wxv = g("coord_mlp_r_virtual.2.weight")[0]
        self.WX, self.bX = g("coord_mlp_v_virtual.0.weight"), g("coord_mlp_v_virtual.0.bias")
        self.wX = g("coord_mlp_v_virtual.2.weight")[0]
        self.W5, self.b5 = g("coord_mlp_vel.0.weight"), g("coord_mlp_vel.0.bias")
        self.w6, self.b6 = g("coord_mlp_vel.2.weight")[0], g("coord_mlp_vel.2.bias")
        if gravity:
            self.Wg, self.bg = g("gravity_mlp.0.weight"), g("gravity_mlp.0.bias")
            self.wg2, self.bg2 = g("gravity_mlp.2.weight")[0], g("gravity_mlp.2.bias")
        U1 = g("node_mlp.0.weight")                      # [H, 2H + H*C]  (:90); u flattened as k*C+c (:157)
        self.U1h, self.U1a = U1[:, :H], U1[:, H:2 * H]
        self.U1u = U1[:, 2 * H:].reshape(H, H, C).permute(2, 0, 1)      # [C][n][k]
        self.e1 = g("node_mlp.0.bias")
        self.U2, self.e2 = g("node_mlp.2.weight"), g("node_mlp.2.bias")
        T1 = g("node_mlp_virtual.0.weight")              # [H, 2H]        (:96)
        self.T1s, self.T1a, self.f1 = T1[:, :H], T1[:, H:], g("node_mlp_virtual.0.bias")
        self.T2, self.f2 = g("node_mlp_virtual.2.weight"), g("node_mlp_virtual.2.bias")
        if attention:
            self.wa, self.ba = g("att_mlp.0.weight")[0], g("att_mlp.0.bias")
            self.wav, self.bav = g("att_mlp_virtual.0.weight")[0], g("att_mlp_virtual.0.bias")


class Graph:
    def __init__(self, row: T, col: T, batch: T, n_graphs: int, n_owned: Optional[int] = None):
        self.row, self.col, self.batch = row, col, batch
        self.N = int(batch.numel()) if n_owned is None else n_owned
        self.B = n_graphs
        deg = torch.bincount(row, minlength=self.N)
        self.dinv = 1.0 / deg.clamp(min=1).double()
        nb = torch.bincount(batch[:self.N], minlength=n_graphs)
        self.inv_nb = 1.0 / nb.clamp(min=1).double()


class Flags:
    def __init__(self, attention=False, normalize=False, tanh=False, gravity=None, eps=1e-8):
        self.attention, self.normalize, self.tanh, self.gravity, self.eps = attention, normalize, tanh, gravity, eps


# =========================================================================== forward phases
def graph_pre(w: LayerWeights, Z, S, xsum, inv_nb):
    """Per graph: xbar, centred Gram M (:212-214) and the per-(graph,channel)
    constant part of phi_ev's first layer, G1 = V1s S_c + V1m M[:,c]."""
    xbar = xsum * inv_nb[:, None]
    Zc = Z - xbar[:, :, None]
    M = torch.einsum("bac,bad->bcd", Zc, Zc)
    G1 = S @ w.V1s.T + torch.einsum("nd,bdc->bcn", w.V1m, M)
    return dict(xbar=xbar, Zc=Zc, M=M, G1=G1)


def node_pre(w: LayerWeights, h, fl: Flags):
    """Per node: the h-dependent halves of every first-layer product + phi_v / phi_g."""
    o = dict(P=h @ w.Ws.T + w.b1, Q=h @ w.Wt.T, Av=h @ w.V1h.T + w.c1, Uh=h @ w.U1h.T + w.e1)
    o["z5"] = h @ w.W5.T + w.b5
    o["sv"] = silu(o["z5"]) @ w.w6 + w.b6
    if fl.gravity is not None:
        o["zg"] = h @ w.Wg.T + w.bg
        o["sg"] = silu(o["zg"]) @ w.wg2 + w.bg2
    return o


def _edge_recompute(w, g: Graph, fl: Flags, P, Q, x, ea):
    row, col = g.row, g.col
    d = x[row] - x[col]
    q = (d * d).sum(1)
    if fl.normalize:
        nrm = q.sqrt() + fl.eps
        dn = d / nrm[:, None]
    else:
        nrm, dn = None, d
    z1 = P[row] + Q[col] + q[:, None] * w.wq + ea @ w.Wa.T
    a1 = silu(z1)
    z2 = a1 @ w.W2.T + w.b2
    m0 = silu(z2)
    if fl.attention:
        gate = torch.sigmoid(m0 @ w.wa + w.ba)
        m = m0 * gate[:, None]
    else:
        gate, m = None, m0
    z3 = m @ w.W3.T + w.b3
    a3 = silu(z3)
    s = a3 @ w.w4
    if fl.tanh:
        s = torch.tanh(s)
    return dict(d=d, q=q, nrm=nrm, dn=dn, z1=z1, a1=a1, z2=z2, m0=m0, gate=gate, m=m, z3=z3, a3=a3, s=s)


def edge_fwd(w, g: Graph, fl: Flags, P, Q, x, ea):
    """Fused real-edge phase (:102-108, :125-129, :156): sums per row of m_e and of d_e*phi_x(m_e)."""
    r = _edge_recompute(w, g, fl, P, Q, x, ea)
    H = P.size(1)
    msum = P.new_zeros(g.N, H).index_add_(0, g.row, r["m"])
    tsum = P.new_zeros(g.N, 3).index_add_(0, g.row, r["dn"] * r["s"][:, None])
    return dict(msum=msum, tsum=tsum)


def _virtual_recompute(w, g: Graph, fl: Flags, Av, G1, x, Z):
    N = g.N
    b = g.batch[:N]
    D = Z[b] - x[:N, :, None]                                  # [N,3,C]   (:206)
    rho = D.norm(dim=1)                                        # [N,C]     (:207)
    zv1 = Av[:, None, :] + G1[b] + rho[:, :, None] * w.vr      # [N,C,H]
    av1 = silu(zv1)
    zv2 = av1 @ w.V2.T + w.c2
    u0 = silu(zv2)
    if fl.attention:
        gv = torch.sigmoid(u0 @ w.wav + w.bav)                 # [N,C]
        u = u0 * gv[:, :, None]
    else:
        gv, u = None, u0
    zxv = u @ w.Wxv.T + w.bxv
    axv = silu(zxv)
    sxv = axv @ w.wxv
    zX = u @ w.WX.T + w.bX
    aX = silu(zX)
    sX = aX @ w.wX
    if fl.tanh:
        sxv, sX = torch.tanh(sxv), torch.tanh(sX)
    return dict(D=D, rho=rho, zv1=zv1, av1=av1, zv2=zv2, u0=u0, gv=gv, u=u, zxv=zxv, axv=axv, sxv=sxv,
                zX=zX, aX=aX, sX=sX)


def virtual_fwd(w, g: Graph, fl: Flags, Av, G1, x, v, Z, tsum, sv, sg):
    """Dense N x C real<->virtual phase (:111-119, :133-150) + new coordinates and
    the per-graph partial sums that the (all-)reduce consumes."""
    N, C = g.N, Z.size(2)
    b = g.batch[:N]
    r = _virtual_recompute(w, g, fl, Av, G1, x, Z)
    xn = x[:N] + tsum * g.dinv[:, None].to(x.dtype) - (r["D"] * r["sxv"][:, None, :]).sum(2) / C + sv[:, None] * v[:N]
    if fl.gravity is not None:
        xn = xn + sg[:, None] * torch.as_tensor(fl.gravity, dtype=x.dtype)
    Dsum = Z.new_zeros(g.B, 3, C).index_add_(0, b, r["D"] * r["sX"][:, None, :])
    Usum = Z.new_zeros(g.B, C, Av.size(1)).index_add_(0, b, r["u"])
    xsum = Z.new_zeros(g.B, 3).index_add_(0, b, xn)
    return dict(x_new=xn, zv2=r["zv2"], u=r["u"], Dsum=Dsum, Usum=Usum, xsum_new=xsum)


def node_h_fwd(w, g: Graph, Uh, msum, u, h):
    """phi_h (:153-166) with the first layer split per source block."""
    zh1 = Uh + (msum * g.dinv[:, None].to(h.dtype)) @ w.U1a.T + torch.einsum("ick,cnk->in", u, w.U1u)
    return dict(zh1=zh1, h_new=h + silu(zh1) @ w.U2.T + w.e2)


def graph_post(w, Z, S, Dsum, Usum, inv_nb):
    """Per graph: Z' (:147-149) and S' (:168-177)."""
    Zn = Z + Dsum * inv_nb[:, None, None].to(Z.dtype)
    zt1 = S @ w.T1s.T + (Usum * inv_nb[:, None, None].to(Z.dtype)) @ w.T1a.T + w.f1
    return dict(Z_new=Zn, zt1=zt1, S_new=S + silu(zt1) @ w.T2.T + w.f2)


# =========================================================================== backward phases
def graph_post_bwd(w, S, Usum, inv_nb, gZn, gSn, last: bool):
    """Returns grads wrt Z (direct), S, Dsum, Usum and phi_hv weight grads."""
    inb = inv_nb[:, None, None].to(S.dtype)
    out = dict(gZ=gZn.clone(), gDsum=gZn * inb, wg={})
    if last:                                   # final S' is discarded (:276): no gradient
        out["gS"] = torch.zeros_like(S)
        out["gUsum"] = torch.zeros_like(Usum)
        return out
    Um = Usum * inb
    zt1 = S @ w.T1s.T + Um @ w.T1a.T + w.f1
    at = silu(zt1)
    gzt2 = gSn
    gzt1 = (gzt2 @ w.T2) * dsilu(zt1)
    f = lambda a: a.reshape(-1, a.size(-1))
    out["wg"] = dict(T2=f(gzt2).T @ f(at), f2=f(gzt2).sum(0), T1s=f(gzt1).T @ f(S), T1a=f(gzt1).T @ f(Um),
                     f1=f(gzt1).sum(0))
    out["gS"] = gSn + gzt1 @ w.T1s
    out["gUsum"] = (gzt1 @ w.T1a) * inb
    return out


def node_h_bwd(w, g: Graph, zh1, msum, u, ghn, last: bool):
    """Backward of phi_h: grads wrt Uh (=gzh1), m-mean (-> gm per edge of the row), u; dh residual."""
    N, H = zh1.shape
    C = u.size(1)
    if last:                                   # final h' is discarded (:276)
        z = torch.zeros_like(zh1)
        return dict(gzh1=z, gm=z.clone(), gu=torch.zeros_like(u), wg={})
    ah = silu(zh1)
    gzh1 = (ghn @ w.U2) * dsilu(zh1)
    mmean = msum * g.dinv[:, None].to(zh1.dtype)
    wg = dict(U2=ghn.T @ ah, e2=ghn.sum(0), U1a=gzh1.T @ mmean, U1u=torch.einsum("in,ick->cnk", gzh1, u))
    gm = (gzh1 @ w.U1a) * g.dinv[:, None].to(zh1.dtype)
    gu = torch.einsum("in,cnk->ick", gzh1, w.U1u)
    return dict(gzh1=gzh1, gm=gm, gu=gu, wg=wg)


def virtual_bwd(w, g: Graph, fl: Flags, Av, G1, x, v, Z, gxn, gDsum, gUsum, gu_h):
    """Backward of the N x C phase.  gxn = dL/dx' [N,3]; gu_h from node_h_bwd."""
    N, C = g.N, Z.size(2)
    b = g.batch[:N]
    r = _virtual_recompute(w, g, fl, Av, G1, x, Z)
    D, rho = r["D"], r["rho"]
    gu = gu_h + gUsum[b]
    gsxv = -(D * gxn[:, :, None]).sum(1) / C                   # [N,C]
    gD = -r["sxv"][:, None, :] * gxn[:, :, None] / C
    gsX = (D * gDsum[b]).sum(1)
    gD = gD + r["sX"][:, None, :] * gDsum[b]
    if fl.tanh:
        gsxv = gsxv * (1 - r["sxv"] ** 2)
        gsX = gsX * (1 - r["sX"] ** 2)
    f = lambda a: a.reshape(-1, a.size(-1))
    gzxv = gsxv[:, :, None] * w.wxv * dsilu(r["zxv"])
    gzX = gsX[:, :, None] * w.wX * dsilu(r["zX"])
    wg = dict(wxv=(gsxv[:, :, None] * r["axv"]).sum((0, 1)), wX=(gsX[:, :, None] * r["aX"]).sum((0, 1)),
              Wxv=f(gzxv).T @ f(r["u"]), bxv=f(gzxv).sum(0), WX=f(gzX).T @ f(r["u"]), bX=f(gzX).sum(0))
    gu = gu + gzxv @ w.Wxv + gzX @ w.WX
    if fl.attention:
        ggv = (gu * r["u0"]).sum(2)
        gu0 = gu * r["gv"][:, :, None]
        gpre = ggv * r["gv"] * (1 - r["gv"])
        wg["wav"] = (gpre[:, :, None] * r["u0"]).sum((0, 1))
        wg["bav"] = gpre.sum().reshape(1)
        gu0 = gu0 + gpre[:, :, None] * w.wav
    else:
        gu0 = gu
    gzv2 = gu0 * dsilu(r["zv2"])
    wg["V2"], wg["c2"] = f(gzv2).T @ f(r["av1"]), f(gzv2).sum(0)
    gzv1 = (gzv2 @ w.V2) * dsilu(r["zv1"])
    gAv = gzv1.sum(1)
    gG1 = torch.zeros_like(G1).index_add_(0, b, gzv1)
    grho = gzv1 @ w.vr
    wg["vr"] = (gzv1 * rho[:, :, None]).sum((0, 1))
    safe = torch.where(rho > 0, rho, torch.ones_like(rho))
    gD = gD + torch.where(rho > 0, grho / safe, torch.zeros_like(rho))[:, None, :] * D
    gx = gxn - gD.sum(2)
    gZ = torch.zeros_like(Z).index_add_(0, b, gD)
    gsv = (gxn * v[:N]).sum(1)
    out = dict(gAv=gAv, gG1=gG1, gx=gx, gZ=gZ, gsv=gsv, gt=gxn * g.dinv[:, None].to(x.dtype), wg=wg)
    if fl.gravity is not None:
        out["gsg"] = gxn @ torch.as_tensor(fl.gravity, dtype=x.dtype)
    return out


def edge_bwd(w, g: Graph, fl: Flags, P, Q, x, ea, gm, gt):
    """Backward of the fused edge phase.  gm [N,H]: grad wrt every m_e of a row;
    gt [N,3]: grad wrt every (dn_e * s_e) of a row.  Returns gP (row-reduced),
    gQ (col-scattered), gx (both ends) and phi_e / phi_x weight grads."""
    r = _edge_recompute(w, g, fl, P, Q, x, ea)
    row, col = g.row, g.col
    Nl = x.size(0)
    gte = gt[row]
    gs = (r["dn"] * gte).sum(1)
    gdn = r["s"][:, None] * gte
    if fl.tanh:
        gs = gs * (1 - r["s"] ** 2)
    gz3 = gs[:, None] * w.w4 * dsilu(r["z3"])
    wg = dict(w4=(gs[:, None] * r["a3"]).sum(0), W3=gz3.T @ r["m"], b3=gz3.sum(0))
    gmm = gm[row] + gz3 @ w.W3
    if fl.attention:
        ggate = (gmm * r["m0"]).sum(1)
        gm0 = gmm * r["gate"][:, None]
        gpre = ggate * r["gate"] * (1 - r["gate"])
        wg["wa"] = (gpre[:, None] * r["m0"]).sum(0)
        wg["ba"] = gpre.sum().reshape(1)
        gm0 = gm0 + gpre[:, None] * w.wa
    else:
        gm0 = gmm
    gz2 = gm0 * dsilu(r["z2"])
    wg["W2"], wg["b2"] = gz2.T @ r["a1"], gz2.sum(0)
    gz1 = (gz2 @ w.W2) * dsilu(r["z1"])
    wg["wq"] = (gz1 * r["q"][:, None]).sum(0)
    wg["Wa"] = gz1.T @ ea
    gq = gz1 @ w.wq
    gd = gdn / r["nrm"][:, None] if fl.normalize else gdn        # norm is detached (:186)
    gd = gd + 2 * gq[:, None] * r["d"]
    H = P.size(1)
    gP = P.new_zeros(g.N, H).index_add_(0, row, gz1)
    gQ = P.new_zeros(Nl, H).index_add_(0, col, gz1)
    gx = x.new_zeros(Nl, 3).index_add_(0, row, gd).index_add_(0, col, -gd)
    return dict(gP=gP, gQ=gQ, gx=gx, wg=wg)


def graph_pre_bwd(w, S, pre, gG1, inv_nb):
    """Backward of graph_pre given the node-summed gG1 [B,C,H]: grads wrt S, Z, xsum."""
    f = lambda a: a.reshape(-1, a.size(-1))
    wg = dict(V1s=f(gG1).T @ f(S), V1m=torch.einsum("bcn,bdc->nd", gG1, pre["M"]))
    gS = gG1 @ w.V1s
    gM = torch.einsum("nd,bcn->bdc", w.V1m, gG1)
    gZc = torch.einsum("bad,bcd->bac", pre["Zc"], gM + gM.transpose(1, 2))
    gxbar = -gZc.sum(2)
    return dict(gS=gS, gZ=gZc, gxsum=gxbar * inv_nb[:, None].to(S.dtype), wg=wg)


def node_pre_bwd(w, g: Graph, fl: Flags, h, gP, gQ, gAv, gUh, gsv, gsg):
    """Backward of node_pre: dh and the weight grads of every h-side block."""
    z5 = h @ w.W5.T + w.b5
    gz5 = gsv[:, None] * w.w6 * dsilu(z5)
    wg = dict(Ws=gP.T @ h, b1=gP.sum(0), Wt=gQ.T @ h, V1h=gAv.T @ h, c1=gAv.sum(0), U1h=gUh.T @ h, e1=gUh.sum(0),
              W5=gz5.T @ h, b5=gz5.sum(0), w6=(gsv[:, None] * silu(z5)).sum(0), b6=gsv.sum().reshape(1))
    gh = gP @ w.Ws + gQ @ w.Wt + gAv @ w.V1h + gUh @ w.U1h + gz5 @ w.W5
    if fl.gravity is not None:
        zg = h @ w.Wg.T + w.bg
        gzg = gsg[:, None] * w.wg2 * dsilu(zg)
        wg.update(Wg=gzg.T @ h, bg=gzg.sum(0), wg2=(gsg[:, None] * silu(zg)).sum(0), bg2=gsg.sum().reshape(1))
        gh = gh + gzg @ w.Wg
    return dict(gh=gh, wg=wg)


# =========================================================================== whole model, staged
def _scatter_wg(grads: Dict[str, T], p: str, wg: Dict[str, T], w: LayerWeights, H: int, C: int, Fe: int):
    """Accumulate phase-level weight grads (kernel names) into reference-named tensors."""
    def acc(key, val):
        grads[key] = grads.get(key, 0) + val
    def put_cols(key, shape, sl, val):
        if key not in grads or isinstance(grads[key], int):
            grads[key] = torch.zeros(shape, dtype=val.dtype)
        grads[key][:, sl] = grads[key][:, sl] + val
    e0 = (H, 2 * H + 1 + Fe)
    v0 = (H, 2 * H + 1 + C)
    n0 = (H, 2 * H + H * C)
    t0 = (H, 2 * H)
    for k, val in wg.items():
        if k == "Ws": put_cols(f"{p}.edge_mlp.0.weight", e0, slice(0, H), val)
        elif k == "Wt": put_cols(f"{p}.edge_mlp.0.weight", e0, slice(H, 2 * H), val)
        elif k == "wq": put_cols(f"{p}.edge_mlp.0.weight", e0, slice(2 * H, 2 * H + 1), val[:, None])
        elif k == "Wa": put_cols(f"{p}.edge_mlp.0.weight", e0, slice(2 * H + 1, 2 * H + 1 + Fe), val)
        elif k == "b1": acc(f"{p}.edge_mlp.0.bias", val)
        elif k == "W2": acc(f"{p}.edge_mlp.2.weight", val)
        elif k == "b2": acc(f"{p}.edge_mlp.2.bias", val)
        elif k == "W3": acc(f"{p}.coord_mlp_r.0.weight", val)
        elif k == "b3": acc(f"{p}.coord_mlp_r.0.bias", val)
        elif k == "w4": acc(f"{p}.coord_mlp_r.2.weight", val[None])
        elif k == "V1h": put_cols(f"{p}.edge_mlp_virtual.0.weight", v0, slice(0, H), val)
        elif k == "V1s": put_cols(f"{p}.edge_mlp_virtual.0.weight", v0, slice(H, 2 * H), val)
        elif k == "vr": put_cols(f"{p}.edge_mlp_virtual.0.weight", v0, slice(2 * H, 2 * H + 1), val[:, None])
        elif k == "V1m": put_cols(f"{p}.edge_mlp_virtual.0.weight", v0, slice(2 * H + 1, 2 * H + 1 + C), val)
        elif k == "c1": acc(f"{p}.edge_mlp_virtual.0.bias", val)
        elif k == "V2": acc(f"{p}.edge_mlp_virtual.2.weight", val)
        elif k == "c2": acc(f"{p}.edge_mlp_virtual.2.bias", val)
        elif k == "Wxv": acc(f"{p}.coord_mlp_r_virtual.0.weight", val)
        elif k == "bxv": acc(f"{p}.coord_mlp_r_virtual.0.bias", val)
        elif k == "wxv": acc(f"{p}.coord_mlp_r_virtual.2.weight", val[None])
        elif k == "WX": acc(f"{p}.coord_mlp_v_virtual.0.weight", val)
        elif k == "bX": acc(f"{p}.coord_mlp_v_virtual.0.bias", val)
        elif k == "wX": acc(f"{p}.coord_mlp_v_virtual.2.weight", val[None])
        elif k == "W5": acc(f"{p}.coord_mlp_vel.0.weight", val)
        elif k == "b5": acc(f"{p}.coord_mlp_vel.0.bias", val)
        elif k == "w6": acc(f"{p}.coord_mlp_vel.2.weight", val[None])
        elif k == "b6": acc(f"{p}.coord_mlp_vel.2.bias", val)
        elif k == "Wg": acc(f"{p}.gravity_mlp.0.weight", val)
        elif k == "bg": acc(f"{p}.gravity_mlp.0.bias", val)
        elif k == "wg2": acc(f"{p}.gravity_mlp.2.weight", val[None])
        elif k == "bg2": acc(f"{p}.gravity_mlp.2.bias", val)
        elif k == "U1h": put_cols(f"{p}.node_mlp.0.weight", n0, slice(0, H), val)
        elif k == "U1a": put_cols(f"{p}.node_mlp.0.weight", n0, slice(H, 2 * H), val)
        elif k == "U1u":        # [C][n][k] -> columns 2H + k*C + c
            put_cols(f"{p}.node_mlp.0.weight", n0, slice(2 * H, 2 * H + H * C), val.permute(1, 2, 0).reshape(H, H * C))
        elif k == "e1": acc(f"{p}.node_mlp.0.bias", val)
        elif k == "U2": acc(f"{p}.node_mlp.2.weight", val)
        elif k == "e2": acc(f"{p}.node_mlp.2.bias", val)
        elif k == "T1s": put_cols(f"{p}.node_mlp_virtual.0.weight", t0, slice(0, H), val)
        elif k == "T1a": put_cols(f"{p}.node_mlp_virtual.0.weight", t0, slice(H, 2 * H), val)
        elif k == "f1": acc(f"{p}.node_mlp_virtual.0.bias", val)
        elif k == "T2": acc(f"{p}.node_mlp_virtual.2.weight", val)
        elif k == "f2": acc(f"{p}.node_mlp_virtual.2.bias", val)
        elif k == "wa": acc(f"{p}.att_mlp.0.weight", val[None])
        elif k == "ba": acc(f"{p}.att_mlp.0.bias", val)
        elif k == "wav": acc(f"{p}.att_mlp_virtual.0.weight", val[None])
        elif k == "bav": acc(f"{p}.att_mlp_virtual.0.bias", val)
        else:
            raise KeyError(k)


class StagedModel:
    """The whole L-layer forward/backward through the phase functions above."""

    def __init__(self, sd: Dict[str, T], H, C, Fe, L, fl: Flags):
        self.sd, self.H, self.C, self.Fe, self.L, self.fl = sd, H, C, Fe, L, fl
        self.w = [LayerWeights(sd, f"gcl_{l}", H, C, Fe, fl.attention, fl.gravity is not None) for l in range(L)]

    def forward(self, node_feat, x0, v, edge_index, batch, loc_mean, edge_attr):
        """edge_index in ORIGINAL order; sorted here exactly as graph-prep does."""
        N = x0.size(0)
        B = int(batch[-1]) + 1
        perm = torch.sort(edge_index[0], stable=True).indices
        self.g = g = Graph(edge_index[0][perm], edge_index[1][perm], batch, B)
        self.ea = ea = edge_attr[perm]
        self.v, self.node_feat = v, node_feat
        dt = x0.dtype
        g.dinv, g.inv_nb = g.dinv.to(dt), g.inv_nb.to(dt)
        h = node_feat @ self.sd["embedding_in.weight"].T + self.sd["embedding_in.bias"]
        S = self.sd["virtual_node_feat"][0].T.unsqueeze(0).repeat(B, 1, 1)          # [B,C,H]
        Z, x = loc_mean, x0
        xsum = x.new_zeros(B, 3).index_add_(0, batch, x)
        self.saved = []
        for l in range(self.L):
            w, last = self.w[l], l == self.L - 1
            pre = graph_pre(w, Z, S, xsum, g.inv_nb)
            npre = node_pre(w, h, self.fl)
            e = edge_fwd(w, g, self.fl, npre["P"], npre["Q"], x, ea)
            vf = virtual_fwd(w, g, self.fl, npre["Av"], pre["G1"], x, v, Z, e["tsum"], npre["sv"], npre.get("sg"))
            sv = dict(h=h, x=x, Z=Z, S=S, pre=pre, npre=npre, e=e, vf=vf)
            if not last:
                nh = node_h_fwd(w, g, npre["Uh"], e["msum"], vf["u"], h)
                sv["nh"] = nh
            gp = graph_post(w, Z, S, vf["Dsum"], vf["Usum"], g.inv_nb)
            sv["gp"] = gp
            self.saved.append(sv)
            x, Z, xsum = vf["x_new"], gp["Z_new"], vf["xsum_new"]
            if not last:
                h, S = nh["h_new"], gp["S_new"]
        return x, Z

    def backward(self, gx_out, gZ_out):
        g, fl, H, C, Fe = self.g, self.fl, self.H, self.C, self.Fe
        grads: Dict[str, T] = {}
        N = g.N
        gh = gx_out.new_zeros(N, H)
        gS = gx_out.new_zeros(g.B, C, H)
        gx, gZ = gx_out, gZ_out
        gxsum_next = gx_out.new_zeros(g.B, 3)
        self.bsaved = [None] * self.L          # per-layer backward intermediates, for phase-level GPU tests
        for l in reversed(range(self.L)):
            w, sv, last, p = self.w[l], self.saved[l], l == self.L - 1, f"gcl_{l}"
            gxn = gx + gxsum_next[g.batch]            # x' also feeds the next layer's xbar
            a = graph_post_bwd(w, sv["S"], sv["vf"]["Usum"], g.inv_nb, gZ, gS, last)
            if last:
                b = node_h_bwd(w, g, sv["npre"]["Uh"], sv["e"]["msum"], sv["vf"]["u"], gh, True)
            else:
                b = node_h_bwd(w, g, sv["nh"]["zh1"], sv["e"]["msum"], sv["vf"]["u"], gh, False)
            c = virtual_bwd(w, g, fl, sv["npre"]["Av"], sv["pre"]["G1"], sv["x"], self.v, sv["Z"], gxn, a["gDsum"],
                            a["gUsum"], b["gu"])
            d = edge_bwd(w, g, fl, sv["npre"]["P"], sv["npre"]["Q"], sv["x"], self.ea, b["gm"], c["gt"])
            e = graph_pre_bwd(w, sv["S"], sv["pre"], c["gG1"], g.inv_nb)
            f = node_pre_bwd(w, g, fl, sv["h"], d["gP"], d["gQ"], c["gAv"], b["gzh1"], c["gsv"], c.get("gsg"))
            if last:                                  # phi_h is dead in the last layer: its tensors get no grad
                f["wg"].pop("U1h"), f["wg"].pop("e1")
            lg: Dict[str, T] = {}
            for wg in (a["wg"], b["wg"], c["wg"], d["wg"], e["wg"], f["wg"]):
                _scatter_wg(grads, p, wg, w, H, C, Fe)
                _scatter_wg(lg, p, wg, w, H, C, Fe)
            self.bsaved[l] = dict(gh_new=gh, gx_new=gx, gZ_new=gZ, gS_new=gS, gxsum_next=gxsum_next, gxn=gxn,
                                  a=a, b=b, c=c, d=d, e=e, f=f, layer_grads=lg)
            gh = gh + f["gh"]
            gx = c["gx"] + d["gx"]
            gZ = a["gZ"] + c["gZ"] + e["gZ"]
            gS = a["gS"] + e["gS"]
            gxsum_next = e["gxsum"]
        gx = gx + gxsum_next[g.batch]                  # layer 0's xbar comes from the input coordinates
        grads["embedding_in.weight"] = gh.T @ self.node_feat
        grads["embedding_in.bias"] = gh.sum(0)
        grads["virtual_node_feat"] = gS.sum(0).T.unsqueeze(0)          # [1,H,C]
        gin = dict(node_loc=gx, loc_mean=gZ, node_feat=gh @ self.sd["embedding_in.weight"])
        return grads, gin
