"""Every CUDA phase (one C-ABI call each) against the matching function of oracle/staged.py,
which tests/test_staged_spec.py proves equal to the oracle's autograd.  A failure here names
the kernel.  Inputs of each phase are the fp64 staged intermediates rounded to fp32."""
import ctypes as C

import pytest
import torch

from oracle import staged
from tests.gpu_util import make_graph_case, rel_err

pytestmark = pytest.mark.gpu

TOL = 3e-5       # relative to the largest entry of each tensor
TOL_W = 2e-4     # weight gradients (long sums)

PHASE_CASES = {
    "c3": dict(seed=31, sizes=[150, 130], deg=11, C=3, L=2),
    "c3_gravity_heavy": dict(seed=32, sizes=[300], deg=14, C=3, L=2, gravity=[0, -1, 0], heavy_row=300),
    "c8": dict(seed=33, sizes=[70, 90], deg=7, C=8, L=2),
    "flags_c2": dict(seed=34, sizes=[100, 120], deg=8, C=2, L=2, attention=True, normalize=True, tanh=True),
    "small_graphs": dict(seed=35, sizes=[5] * 60, deg=2, C=3, L=2),
}


def _setup(name):
    from fastegnn_b200 import _lib as L
    from fastegnn_b200.ops import CsrGraph, SavedBlock, layer_ptrs, make_dims
    cfg, params, inp = make_graph_case(**PHASE_CASES[name])
    p64 = {k: v.double() for k, v in params.items()}
    i64 = {k: (v.double() if v is not None and v.is_floating_point() else v) for k, v in inp.items()}
    fl = staged.Flags(cfg.attention, cfg.normalize, cfg.tanh, cfg.gravity, cfg.eps)
    sm = staged.StagedModel(p64, 64, cfg.virtual_channels, cfg.edge_attr_nf, cfg.n_layers, fl)
    sm.forward(i64["node_feat"], i64["node_loc"], i64["node_vel"], i64["edge_index"], i64["data_batch"],
               i64["loc_mean"], i64["edge_attr"])
    sm.backward(i64["wx"], i64["wz"])
    dev = torch.device("cuda:0")
    B = len(PHASE_CASES[name]["sizes"])
    graph = CsrGraph(inp["edge_index"].to(dev), inp["data_batch"].to(dev), inp["edge_attr"].to(dev), B)
    flags = (L.F_ATTENTION if cfg.attention else 0) | (L.F_NORMALIZE if cfg.normalize else 0) | \
            (L.F_TANH if cfg.tanh else 0) | (L.F_GRAVITY if cfg.gravity is not None else 0)
    gparams = {k: v.to(dev).contiguous() for k, v in params.items()}
    return dict(cfg=cfg, sm=sm, dev=dev, graph=graph, flags=flags, gparams=gparams, L=L, v=inp["node_vel"].to(dev),
                make_dims=make_dims, SavedBlock=SavedBlock, layer_ptrs=layer_ptrs)


def _g(t, dev):
    return t.float().contiguous().to(dev)


def _chk(errs, name, got, want, tol):
    e = rel_err(got.detach().cpu(), want)
    errs.append((name, e, tol))


@pytest.mark.parametrize("name", list(PHASE_CASES))
@pytest.mark.parametrize("layer", [0, 1])
def test_phases(name, layer):
    """Every phase on the fp32 FMA kernels (the tensor-core modes of the edge phase have their own tests below)."""
    from fastegnn_b200 import _lib
    _lib.set_precision("fp32")
    try:
        _phases(name, layer)
    finally:
        _lib.set_precision("tf32")


NODE_BWD_TC_TOL = 4e-3      # tcgen05 TF32 tiles of node_pre_backward / node_h_backward (outputs and their weight gradients)


@pytest.mark.parametrize("name", list(PHASE_CASES))
@pytest.mark.parametrize("layer", [0, 1])
def test_node_backward_tensor_core_mode(name, layer):
    """fegnn_node_pre_backward / fegnn_node_h_backward on the tcgen05 TF32 kernel (dense_tc.cu, "node_backward" mode 1),
    every other phase on the fp32 kernels: the same staged inputs, TF32-grade bound on the two phases' outputs and on
    the weight gradients they produce (first-Linear blocks, phi_h, the phi_v / phi_g heads)."""
    from fastegnn_b200 import _lib
    _lib.set_precision("fp32")
    _lib.set_mode("node_backward", 1)
    try:
        _phases(name, layer, tc_node_bwd=True)
    finally:
        _lib.set_precision("tf32")


def _phases(name, layer, tc_node_bwd=False):
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    l = layer
    last = l == cfg.n_layers - 1
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"] | (L.F_LAST if last else 0), cfg.gravity)
    pd, pg = C.byref(dims), C.byref(graph.c)
    prefix = f"gcl_{l}"
    ptrs = s["layer_ptrs"](s["gparams"], prefix)
    pp = C.byref(ptrs)
    sv = s["SavedBlock"](dims, dev)
    sv.buf.zero_()
    ps = C.byref(sv.c)
    S_ = sm.saved[l]
    h, x, Z, Sx = _g(S_["h"], dev), _g(S_["x"], dev), _g(S_["Z"], dev), _g(S_["S"], dev)
    v = s["v"]
    xsum = _g(torch.zeros(B, 3, dtype=torch.float64).index_add_(0, sm.g.batch, S_["x"]), dev)
    errs = []
    e32 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)

    # ---------------- forward phases (each fed with the staged inputs)
    xs2 = e32(B, 3)
    L.check(lib.fegnn_graph_xsum(N, B, L.ptr(x), L.ptr(graph.batch), L.ptr(xs2), st))
    _chk(errs, "graph_xsum", xs2, xsum.cpu().double(), TOL)
    L.check(lib.fegnn_graph_pre_forward(pd, pg, pp, L.ptr(Z), L.ptr(Sx), L.ptr(xsum), ps, st))
    for k, shp in (("M", (B, Cc, Cc)), ("Zc", (B, 3, Cc)), ("G1", (B, Cc, H))):
        _chk(errs, "graph_pre." + k, sv.view(k, shp), S_["pre"][k], TOL)
    L.check(lib.fegnn_node_pre_forward(pd, pp, L.ptr(h), ps, st))
    names = ["P", "Q", "Av", "sv"] + ([] if last else ["Uh"]) + (["sg"] if cfg.gravity is not None else [])
    for k in names:
        shp = (N,) if k in ("sv", "sg") else (N, H)
        _chk(errs, "node_pre." + k, sv.view(k, shp), S_["npre"][k], TOL)
    # edge phase from exact P, Q
    sv.view("P", (N, H)).copy_(_g(S_["npre"]["P"], dev))
    sv.view("Q", (N, H)).copy_(_g(S_["npre"]["Q"], dev))
    L.check(lib.fegnn_edge_forward(pd, pg, pp, L.ptr(x), ps, st))
    _chk(errs, "edge.msum", sv.view("msum", (N, H)), S_["e"]["msum"], TOL)
    _chk(errs, "edge.tsum", sv.view("tsum", (N, 3)), S_["e"]["tsum"], TOL)
    # virtual phase from exact Av, G1, tsum, sv, sg
    sv.view("Av", (N, H)).copy_(_g(S_["npre"]["Av"], dev))
    sv.view("G1", (B, Cc, H)).copy_(_g(S_["pre"]["G1"], dev))
    sv.view("tsum", (N, 3)).copy_(_g(S_["e"]["tsum"], dev))
    sv.view("sv", (N,)).copy_(_g(S_["npre"]["sv"], dev))
    if cfg.gravity is not None:
        sv.view("sg", (N,)).copy_(_g(S_["npre"]["sg"], dev))
    x_new, xsum_new = e32(N, 3), e32(B, 3)
    L.check(lib.fegnn_virtual_forward(pd, pg, pp, L.ptr(x), L.ptr(v), L.ptr(Z), ps, L.ptr(x_new), L.ptr(xsum_new), st))
    _chk(errs, "virtual.x_new", x_new, S_["vf"]["x_new"], TOL)
    _chk(errs, "virtual.u", sv.view("u", (N, Cc, H)), S_["vf"]["u"], TOL)
    _chk(errs, "virtual.Dsum", sv.view("Dsum", (B, 3, Cc)), S_["vf"]["Dsum"], TOL)
    _chk(errs, "virtual.Usum", sv.view("Usum", (B, Cc, H)), S_["vf"]["Usum"], TOL)
    _chk(errs, "virtual.xsum_new", xsum_new, S_["vf"]["xsum_new"], TOL)
    # node_h from exact Uh, msum, u
    sv.view("msum", (N, H)).copy_(_g(S_["e"]["msum"], dev))
    sv.view("u", (N, Cc, H)).copy_(_g(S_["vf"]["u"], dev))
    if not last:
        sv.view("Uh", (N, H)).copy_(_g(S_["npre"]["Uh"], dev))
        h_new = e32(N, H)
        L.check(lib.fegnn_node_h_forward(pd, pg, pp, L.ptr(h), ps, L.ptr(h_new), st))
        _chk(errs, "node_h.zh1", sv.view("zh1", (N, H)), S_["nh"]["zh1"], TOL)
        _chk(errs, "node_h.h_new", h_new, S_["nh"]["h_new"], TOL)
        sv.view("zh1", (N, H)).copy_(_g(S_["nh"]["zh1"], dev))
    sv.view("Dsum", (B, 3, Cc)).copy_(_g(S_["vf"]["Dsum"], dev))
    sv.view("Usum", (B, Cc, H)).copy_(_g(S_["vf"]["Usum"], dev))
    Z_new, S_new = e32(B, 3, Cc), e32(B, Cc, H)
    L.check(lib.fegnn_graph_post_forward(pd, pg, pp, L.ptr(Z), L.ptr(Sx), ps, L.ptr(Z_new), L.ptr(S_new), st))
    _chk(errs, "graph_post.Z_new", Z_new, S_["gp"]["Z_new"], TOL)
    if not last:
        _chk(errs, "graph_post.S_new", S_new, S_["gp"]["S_new"], TOL)
    sv.view("M", (B, Cc, Cc)).copy_(_g(S_["pre"]["M"], dev))
    sv.view("Zc", (B, 3, Cc)).copy_(_g(S_["pre"]["Zc"], dev))

    # ---------------- backward phases
    bs = sm.bsaved[l]
    gviews = {k: torch.zeros_like(p) for k, p in s["gparams"].items() if k.startswith(prefix + ".")}
    gr = s["layer_ptrs"](gviews, prefix)
    pgr = C.byref(gr)
    a, b, c, d_, e_, f_ = bs["a"], bs["b"], bs["c"], bs["d"], bs["e"], bs["f"]
    gZ, gS, gDsum, gUsum = e32(B, 3, Cc), e32(B, Cc, H), e32(B, 3, Cc), e32(B, Cc, H)
    gZn, gSn = _g(bs["gZ_new"], dev), _g(bs["gS_new"], dev)
    L.check(lib.fegnn_graph_post_backward(pd, pg, pp, pgr, L.ptr(Sx), ps, L.ptr(gZn), L.ptr(gSn), L.ptr(gZ), L.ptr(gS),
                                          L.ptr(gDsum), L.ptr(gUsum), st))
    _chk(errs, "graph_post_bwd.gZ", gZ, a["gZ"], TOL)
    _chk(errs, "graph_post_bwd.gS", gS, a["gS"], TOL)
    _chk(errs, "graph_post_bwd.gDsum", gDsum, a["gDsum"], TOL)
    _chk(errs, "graph_post_bwd.gUsum", gUsum, a["gUsum"], TOL)
    gh_new = _g(bs["gh_new"], dev)
    gzh1, gm, gu = e32(N, H), e32(N, H), e32(N, Cc, H)
    if not last:
        L.check(lib.fegnn_node_h_backward(pd, pg, pp, pgr, ps, L.ptr(gh_new), L.ptr(gzh1), L.ptr(gm), L.ptr(gu), st))
        _chk(errs, "node_h_bwd.gzh1", gzh1, b["gzh1"], TOL)
        _chk(errs, "node_h_bwd.gm", gm, b["gm"], TOL)
        _chk(errs, "node_h_bwd.gu", gu, b["gu"], TOL)
    gx_new = _g(bs["gx_new"], dev)
    gxsum_next = _g(bs["gxsum_next"], dev)
    gDsum_x, gUsum_x, gu_x = _g(a["gDsum"], dev), _g(a["gUsum"], dev), _g(b["gu"], dev)
    gAv, gG1, gx = e32(N, H), e32(B, Cc, H), e32(N, 3)
    gsv, gsg, gt = e32(N), e32(N), e32(N, 3)
    gZ_acc = _g(a["gZ"], dev)
    L.check(lib.fegnn_virtual_backward(pd, pg, pp, pgr, L.ptr(x), L.ptr(v), L.ptr(Z), ps, L.ptr(gx_new),
                                       L.ptr(gxsum_next), L.ptr(gDsum_x), None if last else L.ptr(gUsum_x),
                                       None if last else L.ptr(gu_x), L.ptr(gAv), L.ptr(gG1), L.ptr(gx),
                                       L.ptr(gZ_acc), L.ptr(gsv), L.ptr(gsg), L.ptr(gt), st))
    _chk(errs, "virtual_bwd.gAv", gAv, c["gAv"], TOL)
    _chk(errs, "virtual_bwd.gG1", gG1, c["gG1"], TOL)
    _chk(errs, "virtual_bwd.gx", gx, c["gx"], TOL)
    _chk(errs, "virtual_bwd.gZ(acc)", gZ_acc, a["gZ"] + c["gZ"], TOL)
    _chk(errs, "virtual_bwd.gsv", gsv, c["gsv"], TOL)
    _chk(errs, "virtual_bwd.gt", gt, c["gt"], TOL)
    if cfg.gravity is not None:
        _chk(errs, "virtual_bwd.gsg", gsg, c["gsg"], TOL)
    gm_x, gt_x = _g(b["gm"], dev), _g(c["gt"], dev)
    gP, gQ = e32(N, H), e32(N, H)
    gx_acc = _g(c["gx"], dev)
    L.check(lib.fegnn_edge_backward(pd, pg, pp, pgr, L.ptr(x), ps, None if last else L.ptr(gm_x), L.ptr(gt_x),
                                    L.ptr(gP), L.ptr(gQ), L.ptr(gx_acc), st))
    _chk(errs, "edge_bwd.gP", gP, d_["gP"], TOL)
    _chk(errs, "edge_bwd.gQ", gQ, d_["gQ"], TOL)
    _chk(errs, "edge_bwd.gx(acc)", gx_acc, c["gx"] + d_["gx"], TOL)
    gG1_x = _g(c["gG1"], dev)
    gS_acc, gZ_acc2, gxsum = _g(a["gS"], dev), _g(a["gZ"] + c["gZ"], dev), e32(B, 3)
    L.check(lib.fegnn_graph_pre_backward(pd, pg, pp, pgr, L.ptr(Sx), ps, L.ptr(gG1_x), L.ptr(gS_acc), L.ptr(gZ_acc2),
                                         L.ptr(gxsum), st))
    _chk(errs, "graph_pre_bwd.gS(acc)", gS_acc, a["gS"] + e_["gS"], TOL)
    _chk(errs, "graph_pre_bwd.gZ(acc)", gZ_acc2, a["gZ"] + c["gZ"] + e_["gZ"], TOL)
    _chk(errs, "graph_pre_bwd.gxsum", gxsum, e_["gxsum"], TOL)
    gh_io = _g(bs["gh_new"], dev)
    gP_x, gQ_x, gAv_x, gUh_x = _g(d_["gP"], dev), _g(d_["gQ"], dev), _g(c["gAv"], dev), _g(b["gzh1"], dev)
    gsv_x = _g(c["gsv"], dev)
    gsg_x = _g(c["gsg"], dev) if cfg.gravity is not None else None
    L.check(lib.fegnn_node_pre_backward(pd, pp, pgr, L.ptr(h), L.ptr(gP_x), L.ptr(gQ_x), L.ptr(gAv_x),
                                        None if last else L.ptr(gUh_x), L.ptr(gsv_x), L.ptr(gsg_x), L.ptr(gh_io), st))
    _chk(errs, "node_pre_bwd.gh", gh_io, bs["gh_new"] + f_["gh"], TOL)
    torch.cuda.synchronize()
    # weight gradients accumulated by all backward phases of this layer
    for k, want in bs["layer_grads"].items():
        _chk(errs, "wgrad." + k, gviews[k], want, TOL_W)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/phases_{name}_l{l}.txt", "w") as fh:
        for n_, e, t in errs:
            fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    if tc_node_bwd:
        touched = ("node_h_bwd.", "node_pre_bwd.", "wgrad.")
        errs = [(n_, e, NODE_BWD_TC_TOL if n_.startswith(touched) else t) for n_, e, t in errs]
        with open(f"gpurun_out/node_bwd_tc_{name}_l{l}.txt", "w") as fh:
            for n_, e, t in errs:
                if n_.startswith(touched):
                    fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    bad = [(n_, f"{e:.3e}") for n_, e, t in errs if not e <= t]
    assert not bad, bad


# tolerance of the tensor-core modes of the fused edge forward (relative to each tensor's max):
#   3 = 3xTF32 error-compensated tiles -> fp32-grade;  1 = single-pass TF32 (10-bit mantissa operands)
EDGE_MODE_TOL = {0: 3e-5, 3: 3e-5, 1: 4e-3}


@pytest.mark.parametrize("mode", [0, 3, 1])
@pytest.mark.parametrize("name", list(PHASE_CASES))
def test_edge_forward_modes(name, mode):
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"], cfg.gravity)
    ptrs = s["layer_ptrs"](s["gparams"], "gcl_0")
    sv = s["SavedBlock"](dims, dev)
    S_ = sm.saved[0]
    sv.view("P", (N, H)).copy_(_g(S_["npre"]["P"], dev))
    sv.view("Q", (N, H)).copy_(_g(S_["npre"]["Q"], dev))
    x = _g(S_["x"], dev)
    old = L.get_mode("edge_forward")
    try:
        L.set_mode("edge_forward", mode)
        L.check(lib.fegnn_edge_forward(C.byref(dims), C.byref(graph.c), C.byref(ptrs), L.ptr(x), C.byref(sv.c), st))
        torch.cuda.synchronize()
    finally:
        L.set_mode("edge_forward", old)
    e_m = rel_err(sv.view("msum", (N, H)).cpu(), S_["e"]["msum"])
    e_t = rel_err(sv.view("tsum", (N, 3)).cpu(), S_["e"]["tsum"])
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/edge_fwd_mode{mode}_{name}.txt", "w") as fh:
        fh.write(f"mode {mode} {name}: msum {e_m:.3e} tsum {e_t:.3e}\n")
    assert e_m <= EDGE_MODE_TOL[mode] and e_t <= EDGE_MODE_TOL[mode], (e_m, e_t)


EDGE_BWD_TOL = {0: (3e-5, 2e-4), 1: (6e-3, 6e-3), 2: (6e-3, 6e-3), 4: (6e-3, 6e-3), 5: (6e-3, 6e-3), 7: (6e-3, 6e-3), 8: (6e-3, 6e-3)}     # (per-node outputs, weight gradients), relative to tensor max
# modes: 0 fp32 FMA; 1 tcgen05 TF32 (operands in shared memory); 2 / 4 tcgen05 TF32, A operands in tensor memory,
# MN-major weight-gradient operands, 256 / 512 threads per tile; 5 tcgen05 kind::f16 (fp16 operands, one power-of-two scale per
# launch and gradient tensor, same 10-bit mantissa as TF32), two tile streams per CTA; 7 / 8 the same operand format with
# packed-fp16 epilogue arithmetic (edge_tc_bwd4.cu), 256 / 512 threads per tile


@pytest.mark.parametrize("mode", [0, 1, 2, 4, 5, 7, 8])
@pytest.mark.parametrize("name", ["c3", "c3_gravity_heavy", "c8", "small_graphs"])
@pytest.mark.parametrize("layer", [0, 1])
def test_edge_backward_modes(name, mode, layer):
    """fegnn_edge_backward alone (fp32 FMA kernel vs tcgen05 TF32 kernel) against staged.edge_bwd."""
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    l = layer
    last = l == cfg.n_layers - 1
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"] | (L.F_LAST if last else 0), cfg.gravity)
    prefix = f"gcl_{l}"
    ptrs = s["layer_ptrs"](s["gparams"], prefix)
    sv = s["SavedBlock"](dims, dev)
    S_, bs = sm.saved[l], sm.bsaved[l]
    sv.view("P", (N, H)).copy_(_g(S_["npre"]["P"], dev))
    sv.view("Q", (N, H)).copy_(_g(S_["npre"]["Q"], dev))
    x = _g(S_["x"], dev)
    gviews = {k: torch.zeros_like(p) for k, p in s["gparams"].items() if k.startswith(prefix + ".")}
    gr = s["layer_ptrs"](gviews, prefix)
    b, c, d_ = bs["b"], bs["c"], bs["d"]
    gm_x, gt_x = _g(b["gm"], dev), _g(c["gt"], dev)
    gP = torch.empty(N, H, device=dev)
    gQ = torch.empty(N, H, device=dev)
    gx_acc = torch.zeros(N, 3, device=dev)
    old = L.get_mode("edge_backward")
    try:
        L.set_mode("edge_backward", mode)
        L.check(lib.fegnn_edge_backward(C.byref(dims), C.byref(graph.c), C.byref(ptrs), C.byref(gr), L.ptr(x),
                                        C.byref(sv.c), None if last else L.ptr(gm_x), L.ptr(gt_x), L.ptr(gP),
                                        L.ptr(gQ), L.ptr(gx_acc), st))
        torch.cuda.synchronize()
    finally:
        L.set_mode("edge_backward", old)
    tol, tol_w = EDGE_BWD_TOL[mode]
    errs = []
    _chk(errs, "gP", gP, d_["gP"], tol)
    _chk(errs, "gQ", gQ, d_["gQ"], tol)
    _chk(errs, "gx", gx_acc, d_["gx"], tol)
    want = {}
    staged._scatter_wg(want, prefix, d_["wg"], sm.w[l], 64, Cc, cfg.edge_attr_nf)
    for k, wv in want.items():
        _chk(errs, "wgrad." + k, gviews[k], wv, tol_w)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/edge_bwd_mode{mode}_{name}_l{l}.txt", "w") as fh:
        for n_, e, t in errs:
            fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    bad = [(n_, f"{e:.3e}") for n_, e, t in errs if not e <= t]
    assert not bad, bad


VIRT_TOL = {0: 3e-5, 1: 4e-3}      # fp32 FMA kernels / tcgen05 TF32 tiles, relative to each tensor's max


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["c3", "c3_gravity_heavy", "c8", "small_graphs"])
@pytest.mark.parametrize("layer", [0, 1])
def test_virtual_forward_modes(name, mode, layer):
    """fegnn_virtual_forward alone (fp32 FMA kernel vs tcgen05 TF32 kernel) against staged.virtual_fwd."""
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    l = layer
    last = l == cfg.n_layers - 1
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"] | (L.F_LAST if last else 0), cfg.gravity)
    ptrs = s["layer_ptrs"](s["gparams"], f"gcl_{l}")
    sv = s["SavedBlock"](dims, dev)
    sv.buf.zero_()
    S_ = sm.saved[l]
    x, Z, v = _g(S_["x"], dev), _g(S_["Z"], dev), s["v"]
    sv.view("Av", (N, H)).copy_(_g(S_["npre"]["Av"], dev))
    sv.view("G1", (B, Cc, H)).copy_(_g(S_["pre"]["G1"], dev))
    sv.view("tsum", (N, 3)).copy_(_g(S_["e"]["tsum"], dev))
    sv.view("sv", (N,)).copy_(_g(S_["npre"]["sv"], dev))
    if cfg.gravity is not None:
        sv.view("sg", (N,)).copy_(_g(S_["npre"]["sg"], dev))
    x_new, xsum_new = torch.empty(N, 3, device=dev), torch.empty(B, 3, device=dev)
    old = L.get_mode("virtual_forward")
    try:
        L.set_mode("virtual_forward", mode)
        L.check(lib.fegnn_virtual_forward(C.byref(dims), C.byref(graph.c), C.byref(ptrs), L.ptr(x), L.ptr(v), L.ptr(Z),
                                          C.byref(sv.c), L.ptr(x_new), L.ptr(xsum_new), st))
        torch.cuda.synchronize()
    finally:
        L.set_mode("virtual_forward", old)
    tol = VIRT_TOL[mode]
    errs = []
    _chk(errs, "x_new", x_new, S_["vf"]["x_new"], tol)
    _chk(errs, "u", sv.view("u", (N, Cc, H)), S_["vf"]["u"], tol)
    _chk(errs, "Dsum", sv.view("Dsum", (B, 3, Cc)), S_["vf"]["Dsum"], tol)
    _chk(errs, "Usum", sv.view("Usum", (B, Cc, H)), S_["vf"]["Usum"], tol)
    _chk(errs, "xsum_new", xsum_new, S_["vf"]["xsum_new"], tol)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/virtual_fwd_mode{mode}_{name}_l{l}.txt", "w") as fh:
        for n_, e, t in errs:
            fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    bad = [(n_, f"{e:.3e}") for n_, e, t in errs if not e <= t]
    assert not bad, bad


VIRT_BWD_TOL = {0: (3e-5, 2e-4), 1: (6e-3, 6e-3)}


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["c3", "c3_gravity_heavy", "c8", "small_graphs"])
@pytest.mark.parametrize("layer", [0, 1])
def test_virtual_backward_modes(name, mode, layer):
    """fegnn_virtual_backward alone (fp32 FMA kernel vs the two tcgen05 TF32 kernels) against staged.virtual_bwd."""
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    l = layer
    last = l == cfg.n_layers - 1
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"] | (L.F_LAST if last else 0), cfg.gravity)
    prefix = f"gcl_{l}"
    ptrs = s["layer_ptrs"](s["gparams"], prefix)
    sv = s["SavedBlock"](dims, dev)
    sv.buf.zero_()
    S_, bs = sm.saved[l], sm.bsaved[l]
    x, Z, v = _g(S_["x"], dev), _g(S_["Z"], dev), s["v"]
    sv.view("Av", (N, H)).copy_(_g(S_["npre"]["Av"], dev))
    sv.view("G1", (B, Cc, H)).copy_(_g(S_["pre"]["G1"], dev))
    sv.view("u", (N, Cc, H)).copy_(_g(S_["vf"]["u"], dev))
    sv.view("sv", (N,)).copy_(_g(S_["npre"]["sv"], dev))
    if cfg.gravity is not None:
        sv.view("sg", (N,)).copy_(_g(S_["npre"]["sg"], dev))
    gviews = {k: torch.zeros_like(p) for k, p in s["gparams"].items() if k.startswith(prefix + ".")}
    gr = s["layer_ptrs"](gviews, prefix)
    a, b, c = bs["a"], bs["b"], bs["c"]
    gx_new, gxsum_next = _g(bs["gx_new"], dev), _g(bs["gxsum_next"], dev)
    gDsum_x, gUsum_x = _g(a["gDsum"], dev), _g(a["gUsum"], dev)
    gu_x = torch.full((N, Cc, H), float("nan"), device=dev) if last else _g(b["gu"], dev)   # last layer: scratch only
    e32 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.float32)
    gAv, gG1, gx = e32(N, H), e32(B, Cc, H), e32(N, 3)
    gsv, gsg, gt = e32(N), e32(N), e32(N, 3)
    gZ_acc = _g(a["gZ"], dev)
    old = L.get_mode("virtual_backward")
    try:
        L.set_mode("virtual_backward", mode)
        L.check(lib.fegnn_virtual_backward(C.byref(dims), C.byref(graph.c), C.byref(ptrs), C.byref(gr), L.ptr(x), L.ptr(v),
                                           L.ptr(Z), C.byref(sv.c), L.ptr(gx_new), L.ptr(gxsum_next), L.ptr(gDsum_x),
                                           None if last else L.ptr(gUsum_x), L.ptr(gu_x), L.ptr(gAv), L.ptr(gG1),
                                           L.ptr(gx), L.ptr(gZ_acc), L.ptr(gsv), L.ptr(gsg), L.ptr(gt), st))
        torch.cuda.synchronize()
    finally:
        L.set_mode("virtual_backward", old)
    tol, tol_w = VIRT_BWD_TOL[mode]
    errs = []
    _chk(errs, "gAv", gAv, c["gAv"], tol)
    _chk(errs, "gG1", gG1, c["gG1"], tol)
    _chk(errs, "gx", gx, c["gx"], tol)
    _chk(errs, "gZ(acc)", gZ_acc, a["gZ"] + c["gZ"], tol)
    _chk(errs, "gsv", gsv, c["gsv"], tol)
    _chk(errs, "gt", gt, c["gt"], tol)
    if cfg.gravity is not None:
        _chk(errs, "gsg", gsg, c["gsg"], tol)
    want = {}
    staged._scatter_wg(want, prefix, c["wg"], sm.w[l], 64, Cc, cfg.edge_attr_nf)
    for k, wv in want.items():
        _chk(errs, "wgrad." + k, gviews[k], wv, tol_w)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/virtual_bwd_mode{mode}_{name}_l{l}.txt", "w") as fh:
        for n_, e, t in errs:
            fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    bad = [(n_, f"{e:.3e}") for n_, e, t in errs if not e <= t]
    assert not bad, bad


NODE_H_FWD_TOL = {0: TOL, 3: 1.2e-5}   # fp32 FMA kernels ; tcgen05 error-compensated 3xTF32 (fp32-grade; 2x the worst observed, 5.9e-6 at K = 576)


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("name", ["c3", "c3_gravity_heavy", "c8", "flags_c2", "small_graphs"])
def test_node_h_forward_modes(name, mode):
    """fegnn_node_h_forward alone (phi_h, models/FastEGNN.py:153-166): the fp32 FMA kernels against the tcgen05 3xTF32 kernel
    (node_tc.cu, "node_forward" mode 3 = default), both against staged.node_h_fwd from exact Uh, msum, u."""
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    l = 0
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"], cfg.gravity)
    ptrs = s["layer_ptrs"](s["gparams"], f"gcl_{l}")
    sv = s["SavedBlock"](dims, dev)
    sv.buf.zero_()
    S_ = sm.saved[l]
    h = _g(S_["h"], dev)
    sv.view("msum", (N, H)).copy_(_g(S_["e"]["msum"], dev))
    sv.view("u", (N, Cc, H)).copy_(_g(S_["vf"]["u"], dev))
    sv.view("Uh", (N, H)).copy_(_g(S_["npre"]["Uh"], dev))
    h_new = torch.full((N, H), float("nan"), device=dev, dtype=torch.float32)
    old = L.get_mode("node_forward")
    try:
        L.set_mode("node_forward", mode)
        L.check(lib.fegnn_node_h_forward(C.byref(dims), C.byref(graph.c), C.byref(ptrs), L.ptr(h), C.byref(sv.c),
                                         L.ptr(h_new), st))
        torch.cuda.synchronize()
    finally:
        L.set_mode("node_forward", old)
    tol = NODE_H_FWD_TOL[mode]
    errs = []
    _chk(errs, "node_h.zh1", sv.view("zh1", (N, H)), S_["nh"]["zh1"], tol)
    _chk(errs, "node_h.h_new", h_new, S_["nh"]["h_new"], tol)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/node_h_fwd_mode{mode}_{name}.txt", "w") as fh:
        for n_, e, t in errs:
            fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    bad = [(n_, f"{e:.3e}") for n_, e, t in errs if not e <= t]
    assert not bad, bad


NODE_PRE_FWD_TOL = {3: 4e-6}    # tcgen05 error-compensated 3xTF32 (fp32-grade; K = 64)


@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("name", ["c3", "c3_gravity_heavy", "c8", "small_graphs"])
@pytest.mark.parametrize("layer", [0, 1])
def test_node_pre_forward_modes(name, mode, layer):
    """fegnn_node_pre_forward alone (fp32 FMA kernel, tcgen05 single-pass TF32 kernel, tcgen05 3xTF32 kernel = default)
    against staged.node_pre."""
    s = _setup(name)
    L, lib = s["L"], s["L"].lib
    cfg, sm, dev, graph = s["cfg"], s["sm"], s["dev"], s["graph"]
    st = torch.cuda.current_stream().cuda_stream
    l = layer
    last = l == cfg.n_layers - 1
    Cc, N, B, H = cfg.virtual_channels, graph.N, graph.B, 64
    dims = s["make_dims"](N, N, graph.E, B, Cc, graph.Fe, s["flags"] | (L.F_LAST if last else 0), cfg.gravity)
    ptrs = s["layer_ptrs"](s["gparams"], f"gcl_{l}")
    sv = s["SavedBlock"](dims, dev)
    sv.buf.fill_(float("nan"))
    S_ = sm.saved[l]
    h = _g(S_["h"], dev)
    old = L.get_mode("node_forward")
    try:
        L.set_mode("node_forward", mode)
        L.check(lib.fegnn_node_pre_forward(C.byref(dims), C.byref(ptrs), L.ptr(h), C.byref(sv.c), st))
        torch.cuda.synchronize()
    finally:
        L.set_mode("node_forward", old)
    tol = NODE_PRE_FWD_TOL.get(mode) or VIRT_TOL[mode]
    errs = []
    names = ["P", "Q", "Av", "sv"] + ([] if last else ["Uh"]) + (["sg"] if cfg.gravity is not None else [])
    for k in names:
        shp = (N,) if k in ("sv", "sg") else (N, H)
        _chk(errs, k, sv.view(k, shp), S_["npre"][k], tol)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/node_pre_fwd_mode{mode}_{name}_l{l}.txt", "w") as fh:
        for n_, e, t in errs:
            fh.write(f"{'FAIL' if not e <= t else 'ok  '} {n_}: {e:.3e} (tol {t:.0e})\n")
    bad = [(n_, f"{e:.3e}") for n_, e, t in errs if not e <= t]
    assert not bad, bad
