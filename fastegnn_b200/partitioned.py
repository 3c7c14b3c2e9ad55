"""Spatially partitioned FastEGNN for ONE large graph across the GPUs of a box (SURVEY.md 5.1 mode 2).

The reference has no distributed code; this is the multi-GPU form of the same layer
(models/FastEGNN.py:192-223).  Nodes are split into 1-D slabs along the longest axis (equal node
counts); a node is owned by one rank, an edge (row=i, col=j) by the owner of i -- the end every
reference aggregation reduces over (:127-129,:156).  Per layer:

  forward   node_pre -> HALO EXCHANGE of (Q_j, x_j) for remote cols j -> edge phase -> virtual phase
            -> ALL-REDUCE of the per-graph partial sums (Dsum 3C, Usum CH, xsum 3 floats per graph)
            -> node_h, graph_post (replicated)
  backward  graph_post_bwd (replicated) -> node_h_bwd -> virtual_bwd -> edge_bwd
            -> REVERSE HALO (remote dQ_j, dx_j summed at the owner) + ALL-REDUCE of (dG1, dZ) partials
            -> graph_pre_bwd (replicated) -> node_pre_bwd
  per step  one all-reduce of the flat weight-gradient buffer.

Collectives go through torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests of the
host logic); all arithmetic is the same C-ABI phase calls as on one GPU (layer_fn.LayerPhases).
The total loss is the SUM over ranks of rank-local losses: terms that depend only on the replicated
virtual coordinates Z must be added on one rank (or divided by the world size).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .layer_fn import LayerPhases
from .ops import CsrGraph, _require_cuda, _stream, layer_ptrs, make_dims

lib = L.lib


# ------------------------------------------------------------------------------------- host-side plan
class SlabPlan:
    """Index bookkeeping of the slab partition (numpy, host).  Pure integer logic -- unit-tested on CPU."""

    def __init__(self, x: np.ndarray, edge_index: np.ndarray, world: int):
        x = np.asarray(x)
        ei = np.asarray(edge_index)
        N = x.shape[0]
        self.world, self.N = world, N
        self.axis = int(np.argmax(x.max(0) - x.min(0)))
        self.order = np.argsort(x[:, self.axis], kind="stable")            # global ids in slab order
        self.bounds = np.array([N * k // world for k in range(world + 1)], dtype=np.int64)
        self.owner = np.empty(N, dtype=np.int32)
        self.local_id = np.empty(N, dtype=np.int64)
        for k in range(world):
            ids = self.order[self.bounds[k]:self.bounds[k + 1]]
            self.owner[ids] = k
            self.local_id[ids] = np.arange(ids.size)
        row, col = ei[0], ei[1]
        self.edge_owner = self.owner[row]
        self.parts: List[Dict[str, np.ndarray]] = []
        halos = []
        for k in range(world):
            eidx = np.nonzero(self.edge_owner == k)[0]                      # keeps the caller's edge order
            r, c = row[eidx], col[eidx]
            remote = self.owner[c] != k
            hg = np.unique(c[remote])                                       # halo nodes (global ids) ...
            key = self.owner[hg].astype(np.int64) * N + self.local_id[hg]   # ... ordered by (owner rank, owner-local id)
            hg = hg[np.argsort(key, kind="stable")]
            halos.append(hg)
            n_own = int(self.bounds[k + 1] - self.bounds[k])
            halo_pos = {int(g): n_own + j for j, g in enumerate(hg)} if hg.size < 4096 else None
            col_local = self.local_id[c].copy()
            if hg.size:
                lut = np.full(N, -1, dtype=np.int64)
                lut[hg] = n_own + np.arange(hg.size)
                col_local[remote] = lut[c[remote]]
            self.parts.append(dict(edge_ids=eidx, row=self.local_id[r], col=col_local, n_own=n_own,
                                   owned=self.order[self.bounds[k]:self.bounds[k + 1]], halo=hg,
                                   recv_counts=np.bincount(self.owner[hg], minlength=world).astype(np.int64)))
        # what every rank must SEND: for destination d, the nodes of d's halo that this rank owns, in d's halo order
        for k in range(world):
            send_idx, send_counts = [], []
            for d in range(world):
                hg = halos[d]
                mine = hg[self.owner[hg] == k] if hg.size else hg
                send_idx.append(self.local_id[mine])
                send_counts.append(mine.size)
            self.parts[k]["send_idx"] = np.concatenate(send_idx).astype(np.int64) if send_idx else np.zeros(0, np.int64)
            self.parts[k]["send_counts"] = np.array(send_counts, dtype=np.int64)

    def localize(self, rank: int, node_arrays: Dict[str, np.ndarray], edge_arrays: Dict[str, np.ndarray]):
        """Slice global per-node / per-edge arrays for `rank`.  Node arrays get owned rows followed by halo rows."""
        p = self.parts[rank]
        ids = np.concatenate([p["owned"], p["halo"]])
        out = {k: np.ascontiguousarray(np.asarray(v)[ids]) for k, v in node_arrays.items()}
        out.update({k: np.ascontiguousarray(np.asarray(v)[p["edge_ids"]]) for k, v in edge_arrays.items()})
        out["edge_index"] = np.stack([p["row"], p["col"]]).astype(np.int64)
        return out


class HaloComm:
    """Device-side halo exchange + small all-reduces for one rank (hooks of LayerPhases)."""

    def __init__(self, plan: SlabPlan, rank: int, device, group=None):
        p = plan.parts[rank]
        self.N, self.Nl = p["n_own"], p["n_own"] + int(p["halo"].size)
        self.send_idx = torch.from_numpy(p["send_idx"]).to(device)
        self.send_counts = [int(v) for v in p["send_counts"]]
        self.recv_counts = [int(v) for v in p["recv_counts"]]
        self.n_send, self.n_recv = int(sum(self.send_counts)), int(sum(self.recv_counts))
        self.group, self.device = group, device
        self.world = plan.world

    # -- forward: owner -> users
    def exchange(self, Q: torch.Tensor, x: torch.Tensor) -> None:
        """Fill rows [N, Nl) of Q [Nl,64] and x [Nl,3] with the owners' current rows (one all-to-all)."""
        send = torch.cat([Q.index_select(0, self.send_idx), x.index_select(0, self.send_idx)], dim=1)
        recv = torch.empty(self.n_recv, send.size(1), device=self.device, dtype=torch.float32)
        dist.all_to_all_single(recv, send, self.recv_counts, self.send_counts, group=self.group)
        if self.n_recv:
            Q[self.N:] = recv[:, :L.H]
            x[self.N:] = recv[:, L.H:]

    # -- backward: users -> owner, summed
    def reduce_back(self, gQ: torch.Tensor, gx: torch.Tensor) -> None:
        send = torch.cat([gQ[self.N:], gx[self.N:]], dim=1).contiguous()
        recv = torch.empty(self.n_send, send.size(1), device=self.device, dtype=torch.float32)
        dist.all_to_all_single(recv, send, self.send_counts, self.recv_counts, group=self.group)
        if self.n_send:
            gQ.index_add_(0, self.send_idx, recv[:, :L.H])
            gx.index_add_(0, self.send_idx, recv[:, L.H:])

    def allreduce(self, *tensors: torch.Tensor) -> None:
        """Sum small per-graph tensors over ranks in ONE collective (rank order is fixed by the backend)."""
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.all_reduce(flat, group=self.group)
        o = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[o:o + n].view_as(t))
            o += n

    # LayerPhases hooks ---------------------------------------------------------------------------
    def after_node_pre(self, ph: LayerPhases, x: torch.Tensor) -> None:
        self.exchange(ph.saved.view("Q", (self.Nl, L.H)), x)

    def after_virtual(self, ph: LayerPhases, xsum_new: torch.Tensor) -> None:
        d = ph.d
        self.allreduce(ph.saved.view("Dsum", (d.B, 3, d.C)), ph.saved.view("Usum", (d.B, d.C, L.H)), xsum_new)

    def after_edge_backward(self, ph: LayerPhases, gQ, gx, gG1, gZ_part) -> None:
        self.reduce_back(gQ, gx)
        self.allreduce(gG1, gZ_part)


def p2p_address_tables(plan: SlabPlan, rank: int, n_layers: int, nl_max: int, base_q, base_x, base_gq, base_gx,
                       row_bytes_q: int = 4 * L.H, row_bytes_x: int = 12):
    """Destination byte addresses of the peer-memory halo kernels for `rank` (pure integer logic, unit-tested on the CPU
    against a simulated address space).  base_* [world] = every rank's base address of its symmetric Q [L, nl_max, 64],
    x [L, nl_max, 3], gQ [nl_max, 64] and gx [2, nl_max, 3] arrays.  Returns int64 numpy arrays
        fwd_q, fwd_x [L][n_send]   where send entry k (grouped by destination, in the destination's halo order) lands,
        bwd_q [n_recv], bwd_x [2][n_recv]   the owner's row of each of this rank's halo rows (gx has two slots)."""
    parts = plan.parts
    dst_row, dst_rank = [], []
    for d in range(plan.world):
        cnt = int(parts[rank]["send_counts"][d])
        off = int(parts[d]["recv_counts"][:rank].sum())                # d's halo is ordered by (owner rank, owner-local id)
        dst_row.append(parts[d]["n_own"] + off + np.arange(cnt, dtype=np.int64))
        dst_rank.append(np.full(cnt, d, dtype=np.int64))
    dst_row = np.concatenate(dst_row) if dst_row else np.zeros(0, np.int64)
    dst_rank = np.concatenate(dst_rank) if dst_rank else np.zeros(0, np.int64)
    bq, bx = np.asarray(base_q, dtype=np.int64)[dst_rank], np.asarray(base_x, dtype=np.int64)[dst_rank]
    fwd_q = [bq + (l * nl_max + dst_row) * row_bytes_q for l in range(n_layers)]
    fwd_x = [bx + (l * nl_max + dst_row) * row_bytes_x for l in range(n_layers)]
    hg = parts[rank]["halo"]
    own, lid = plan.owner[hg].astype(np.int64), plan.local_id[hg].astype(np.int64)
    bwd_q = np.asarray(base_gq, dtype=np.int64)[own] + lid * row_bytes_q
    bgx = np.asarray(base_gx, dtype=np.int64)[own]
    bwd_x = [bgx + (slot * nl_max + lid) * row_bytes_x for slot in range(2)]
    return fwd_q, fwd_x, bwd_q, bwd_x


class P2PHaloComm(HaloComm):
    """Halo exchange by direct peer-memory access over NVLink / NVSwitch (csrc/halo.cu) instead of NCCL all-to-all.

    Q and x of every layer, and the gradient buffers gQ / gx, live in torch symmetric memory, so every rank can address
    its peers' arrays.  Forward: ONE kernel stores the owners' (Q_j, x_j) rows straight into the users' halo rows, then a
    symmetric-memory barrier on the stream.  Backward: barrier, ONE kernel adds the users' (dQ_j, dx_j) halo rows into the
    owners' rows with remote atomics, barrier.  Destination addresses are fixed by the plan and precomputed as 64-bit
    tables.  The tiny per-graph all-reduces stay on NCCL.  Measured on 2 B200 (tools/p2p_probe.py, 6 000 rows): 20 us per
    exchange against 90 us for index_select + cat + all_to_all_single + two slice copies."""

    def __init__(self, plan: SlabPlan, rank: int, device, n_layers: int, group=None):
        super().__init__(plan, rank, device, group)
        import torch.distributed._symmetric_memory as symm
        grp = dist.group.WORLD if group is None else group
        parts = plan.parts
        self.n_layers = n_layers
        self.Nl_max = Nm = max(int(p["n_own"] + p["halo"].size) for p in parts)
        f32 = dict(dtype=torch.float32, device=device)
        self.Qs, self.xs = symm.empty(n_layers, Nm, L.H, **f32), symm.empty(n_layers, Nm, 3, **f32)
        self.gQs, self.gxs = symm.empty(Nm, L.H, **f32), symm.empty(2, Nm, 3, **f32)
        self.hQ, self.hx = symm.rendezvous(self.Qs, grp), symm.rendezvous(self.xs, grp)
        self.hgQ, self.hgx = symm.rendezvous(self.gQs, grp), symm.rendezvous(self.gxs, grp)
        for t in (self.Qs, self.xs, self.gQs, self.gxs):
            t.zero_()
        i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(device)
        fq, fx, bq, bx = p2p_address_tables(plan, rank, n_layers, Nm, self.hQ.buffer_ptrs, self.hx.buffer_ptrs,
                                            self.hgQ.buffer_ptrs, self.hgx.buffer_ptrs)
        self.send_idx32 = self.send_idx.to(torch.int32)
        self.fwd_q, self.fwd_x = [i64(a) for a in fq], [i64(a) for a in fx]
        self.bwd_q, self.bwd_x = i64(bq), [i64(a) for a in bx]
        self._slot = 0
        self._layer_of = {}

    def barrier(self) -> None:
        self.hQ.barrier(channel=0)

    # -- forward
    def bind_layer(self, ph: LayerPhases, layer: int) -> None:
        """This layer's Q is the symmetric array (node_pre writes the owned rows there, peers write the halo rows)."""
        ph.saved.c.Q = self.Qs[layer].data_ptr()
        self._layer_of[id(ph)] = layer

    def layer_x(self, layer: int, x: torch.Tensor, rows: int) -> torch.Tensor:
        """The layer's input coordinates inside symmetric memory ([Nl,3] view); `rows` leading rows of x are copied."""
        xs = self.xs[layer, :self.Nl]
        xs[:rows].copy_(x[:rows])
        return xs

    def after_node_pre(self, ph: LayerPhases, x: torch.Tensor) -> None:
        l = self._layer_of[id(ph)]
        L.check(lib.fegnn_halo_push(self.n_send, L.ptr(self.send_idx32), L.ptr(self.fwd_q[l]), L.ptr(self.fwd_x[l]),
                                    L.ptr(self.Qs[l]), L.ptr(x), _stream()), "fegnn_halo_push")
        self.barrier()

    # -- backward
    def grad_halo_buffers(self, ph: LayerPhases):
        self._slot ^= 1
        return self.gQs[:self.Nl], self.gxs[self._slot, :self.Nl]

    def after_edge_backward(self, ph: LayerPhases, gQ, gx, gG1, gZ_part) -> None:
        self.barrier()                       # every rank has zero-filled and accumulated its own gQ / gx of this layer
        L.check(lib.fegnn_halo_reduce_push(self.n_recv, self.N, L.ptr(self.bwd_q), L.ptr(self.bwd_x[self._slot]),
                                           L.ptr(gQ), L.ptr(gx), _stream()), "fegnn_halo_reduce_push")
        self.barrier()
        self.allreduce(gG1, gZ_part)


# ------------------------------------------------------------------------------------- autograd driver
class _PartitionedStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner: "PartitionedFastEGNN", graph: CsrGraph, node_feat, x0, v, loc_mean, *params):
        mod, comm = runner.model, runner.comm
        dev = x0.device
        N, Nl, B, Cc, Lyr = comm.N, comm.Nl, graph.B, mod.virtual_channels, mod.n_layers
        st = _stream()
        named = dict(mod.named_parameters())
        h = torch.empty(N, L.H, device=dev, dtype=torch.float32)
        L.check(lib.fegnn_embed_forward(N, node_feat.size(1), L.ptr(node_feat), L.ptr(mod.embedding_in.weight),
                                        L.ptr(mod.embedding_in.bias), L.ptr(h), st), "embed_forward")
        S = mod.virtual_node_feat.detach()[0].t().contiguous().unsqueeze(0).repeat(B, 1, 1).contiguous()
        Z = loc_mean
        p2p = isinstance(comm, P2PHaloComm)
        if p2p:
            comm.barrier()                                                 # peers are done reading last step's arrays
        x = x0.clone()                                                     # [Nl,3]; halo rows refreshed every layer
        xsum = torch.empty(B, 3, device=dev, dtype=torch.float32)
        L.check(lib.fegnn_graph_xsum(N, B, L.ptr(x), L.ptr(graph.batch), L.ptr(xsum), st), "graph_xsum")
        comm.allreduce(xsum)
        phases, states = [], []
        for l in range(Lyr):
            flags = mod._flag_word | (L.F_LAST if l == Lyr - 1 else 0)
            dims = make_dims(N, Nl, graph.E, B, Cc, graph.Fe, flags, mod._gravity)
            ph = LayerPhases(dims, graph, layer_ptrs(named, "gcl_%d" % l), dev)
            if p2p:
                comm.bind_layer(ph, l)
                x = comm.layer_x(l, x, Nl if l == 0 else N)
            states.append((h, x, Z, S))
            h_new, x_new, Z_new, S_new, xsum = ph.forward(h, x, v, Z, S, xsum, hooks=comm)
            phases.append(ph)
            h, x, Z, S = h_new, x_new, Z_new, S_new
        ctx.runner, ctx.graph, ctx.phases, ctx.states = runner, graph, phases, states
        ctx.save_for_backward(node_feat, v)
        ctx.names = list(named.keys())
        return x[:N].contiguous(), Z

    @staticmethod
    def backward(ctx, gx_out, gZ_out):
        runner, graph, phases, states = ctx.runner, ctx.graph, ctx.phases, ctx.states
        mod, comm = runner.model, runner.comm
        node_feat, v = ctx.saved_tensors
        dev = v.device
        N, Nl, B, Cc, Lyr = comm.N, comm.Nl, graph.B, mod.virtual_channels, mod.n_layers
        st = _stream()
        rank0 = runner.rank == 0
        named = dict(mod.named_parameters())
        views = {n: torch.zeros_like(p) for n, p in named.items()}
        null_grads = L.LayerPtrs()                                        # replicated phases: weight grads on rank 0 only
        gZ_new = torch.zeros(B, 3, Cc, device=dev) if gZ_out is None else gZ_out.contiguous().float().clone()
        comm.allreduce(gZ_new)                                            # total dL/dZ (loss = sum of rank-local losses)
        gx_new = torch.zeros(Nl, 3, device=dev)
        if gx_out is not None:
            gx_new[:N] = gx_out
        gh = torch.zeros(N, L.H, device=dev)
        gS_new, gxsum_next = None, None
        for l in range(Lyr - 1, -1, -1):
            ph = phases[l]
            h, x, Z, S = states[l]
            live = {k: t for k, t in views.items() if k not in mod._dead_names}
            grads = layer_ptrs(live, "gcl_%d" % l)
            gh, gx_new, gZ_new, gS_new, gxsum_next = ph.backward(
                grads, h, x, v, Z, S, gh, gx_new, gZ_new, gS_new, gxsum_next, hooks=comm,
                graph_grads=grads if rank0 else null_grads)
        g_x0 = gx_new[:N] + gxsum_next[graph.batch.long()]
        g_nf = torch.empty_like(node_feat) if node_feat.requires_grad else None
        L.check(lib.fegnn_embed_backward(N, node_feat.size(1), L.ptr(node_feat), L.ptr(mod.embedding_in.weight),
                                         L.ptr(gh), L.ptr(views["embedding_in.weight"]),
                                         L.ptr(views["embedding_in.bias"]), L.ptr(g_nf), st), "embed_backward")
        if rank0:
            views["virtual_node_feat"].copy_(gS_new.sum(0).t().unsqueeze(0))
        runner.last_local_grads = views
        grads_out = tuple(None if n in mod._dead_names else views[n] for n in ctx.names)
        g_x0_full = torch.zeros(Nl, 3, device=dev)
        g_x0_full[:N] = g_x0
        return (None, None, g_nf, g_x0_full, None, gZ_new) + grads_out


class PartitionedFastEGNN:
    """Runs a FastEGNN module (same weights on every rank) on this rank's slab of one big graph."""

    def __init__(self, model, plan: SlabPlan, rank: int, device, group=None, halo: str = "nccl"):
        """halo = "nccl": pack + all-to-all + unpack; "p2p": direct peer stores / remote atomics over NVLink
        (P2PHaloComm; needs torch symmetric memory, i.e. all ranks on one NVLink / NVSwitch domain)."""
        self.model, self.plan, self.rank = model, plan, rank
        if halo == "p2p":
            self.comm = P2PHaloComm(plan, rank, device, model.n_layers, group)
        elif halo == "nccl":
            self.comm = HaloComm(plan, rank, device, group)
        else:
            raise ValueError(f"halo must be 'nccl' or 'p2p', got {halo!r}")
        self.last_local_grads = None

    def __call__(self, node_feat, node_loc, node_vel, edge_index, loc_mean, edge_attr, n_global: int):
        """node_* hold this rank's owned rows followed by its halo rows ([Nl, .]); edge_index is local
        (row < N owned, col < Nl); loc_mean [B,3,C] is replicated.  Returns (x' of the OWNED rows, Z')."""
        _require_cuda(node_loc, "node_loc")
        N = self.comm.N
        B = int(loc_mean.size(0))
        batch = torch.zeros(N, dtype=torch.int64, device=node_loc.device)
        graph = CsrGraph(edge_index, batch, edge_attr, B, n_local=self.comm.Nl)
        graph.inv_nb.fill_(1.0 / max(1, n_global))                        # graph size is global, not the slab's
        params = [p for _, p in self.model.named_parameters()]
        return _PartitionedStackFn.apply(self, graph, node_feat[:N].contiguous().float(),
                                         node_loc.contiguous().float(), node_vel[:N].contiguous().float(),
                                         loc_mean.contiguous().float(), *params)

    def allreduce_gradients(self) -> None:
        """Sum the weight gradients over ranks (one collective on a flat buffer)."""
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, group=self.comm.group)
        torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))
