"""Host logic of the partitioned multi-GPU path, on CPU: the slab plan (pure integer bookkeeping) and
the halo / all-reduce plumbing over torch.distributed with the gloo backend at world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _cloud(n=400, deg=8, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3)).astype(np.float32) * np.array([4.0, 1.0, 1.0], dtype=np.float32)
    from scipy.spatial import cKDTree
    r = (deg / n * 4.0 / (4.0 / 3.0 * np.pi)) ** (1.0 / 3.0)
    pairs = cKDTree(x).query_pairs(r, output_type="ndarray")
    ei = np.concatenate([pairs, pairs[:, ::-1]]).T.copy()
    ei = ei[:, rng.permutation(ei.shape[1])]
    return x, ei


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_slab_plan_reassembles_the_global_graph(world):
    from fastegnn_b200.partitioned import SlabPlan
    x, ei = _cloud()
    plan = SlabPlan(x, ei, world)
    N, E = x.shape[0], ei.shape[1]
    assert sorted(np.concatenate([p["owned"] for p in plan.parts]).tolist()) == list(range(N))
    assert sorted(np.concatenate([p["edge_ids"] for p in plan.parts]).tolist()) == list(range(E))
    sizes = [p["n_own"] for p in plan.parts]
    assert max(sizes) - min(sizes) <= 1
    for k, p in enumerate(plan.parts):
        ids = np.concatenate([p["owned"], p["halo"]])                # local id -> global id
        assert (p["row"] < p["n_own"]).all() and (p["col"] < ids.size).all()
        assert np.array_equal(ids[p["row"]], ei[0][p["edge_ids"]])   # local edges map back to the global ones
        assert np.array_equal(ids[p["col"]], ei[1][p["edge_ids"]])
        assert (plan.owner[p["halo"]] != k).all()
        assert np.all(np.diff(plan.owner[p["halo"]]) >= 0)           # halo grouped by owner rank
        assert p["recv_counts"].sum() == p["halo"].size and p["recv_counts"][k] == 0
        # what the others send me is exactly my halo, in my halo order
        got = []
        for src in range(world):
            q = plan.parts[src]
            o = int(q["send_counts"][:k].sum())
            got.append(q["owned"][q["send_idx"][o:o + int(q["send_counts"][k])]])
            assert int(q["send_counts"][k]) == int(p["recv_counts"][src])
        assert np.array_equal(np.concatenate(got), p["halo"])
    # slabs are ordered along the longest axis
    ax = plan.axis
    assert ax == 0
    for k in range(world - 1):
        assert x[plan.parts[k]["owned"], ax].max() <= x[plan.parts[k + 1]["owned"], ax].min()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, x, ei, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastegnn_b200.partitioned import HaloComm, SlabPlan
        plan = SlabPlan(x, ei, world)
        comm = HaloComm(plan, rank, torch.device("cpu"))
        p = plan.parts[rank]
        N, Nl = comm.N, comm.Nl
        ids = np.concatenate([p["owned"], p["halo"]])
        # forward exchange: owners hold f(global id); after the exchange every halo row must hold its owner's value
        gid = torch.from_numpy(ids.astype(np.float32))
        Q = torch.zeros(Nl, 64)
        xx = torch.zeros(Nl, 3)
        Q[:N] = gid[:N, None] * 2 + torch.arange(64)[None]
        xx[:N] = gid[:N, None] * 3 + torch.arange(3)[None]
        comm.exchange(Q, xx)
        ok_fwd = bool(torch.equal(Q, gid[:, None] * 2 + torch.arange(64)[None]) and
                      torch.equal(xx, gid[:, None] * 3 + torch.arange(3)[None]))
        # reverse: every use of a remote node contributes 1 -> owners end with the number of remote users
        gQ = torch.zeros(Nl, 64)
        gx = torch.zeros(Nl, 3)
        gQ[N:] = 1.0
        gx[N:] = 1.0
        comm.reduce_back(gQ, gx)
        users = np.zeros(x.shape[0])
        for q in plan.parts:
            users[q["halo"]] += 1
        ok_bwd = bool(np.array_equal(gQ[:N, 0].numpy(), users[p["owned"]]) and
                      np.array_equal(gx[:N, 2].numpy(), users[p["owned"]]))
        a, b = torch.full((2, 3), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))
        comm.allreduce(a, b)
        tot = world * (world + 1) / 2
        ok_ar = bool((a == tot).all() and (b == 10 * tot).all())
        out[rank] = (ok_fwd, ok_bwd, ok_ar)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_allreduce_over_gloo(world):
    x, ei = _cloud(n=300, deg=7, seed=1)
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), x, ei, out), nprocs=world, join=True)
    assert all(out[r] == (True, True, True) for r in range(world)), dict(out)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_memory_address_tables_in_a_simulated_address_space(world):
    """The 64-bit destination tables of the peer-memory halo kernels (fegnn_halo_push / fegnn_halo_reduce_push): every
    rank's symmetric arrays are numpy buffers at made-up base addresses; executing the tables as the kernels do (row
    stores forward, row adds backward) must reproduce the NCCL-style exchange -- halo rows equal the owners' rows in every
    layer, owners receive the sum of their users' halo gradients -- and never touch a byte outside the target rows."""
    from fastegnn_b200.partitioned import SlabPlan, p2p_address_tables
    x, ei = _cloud(n=500, deg=10, seed=3)
    plan = SlabPlan(x, ei, world)
    Lyr, H = 3, 64
    nl = [p["n_own"] + p["halo"].size for p in plan.parts]
    nm = max(nl)
    rng = np.random.default_rng(1)
    base_q = [(r + 1) << 40 for r in range(world)]
    base_x = [((r + 1) << 40) + (1 << 36) for r in range(world)]
    base_gq = [((r + 1) << 40) + (2 << 36) for r in range(world)]
    base_gx = [((r + 1) << 40) + (3 << 36) for r in range(world)]
    Q = [rng.standard_normal((Lyr, nm, H)).astype(np.float32) for _ in range(world)]
    X = [rng.standard_normal((Lyr, nm, 3)).astype(np.float32) for _ in range(world)]
    Qref = [q.copy() for q in Q]
    Xref = [v.copy() for v in X]

    def locate(addr, bases, row_bytes, rows_total):
        r = int(addr >> 40) - 1
        off = addr - bases[r]
        assert 0 <= off < rows_total * row_bytes and off % row_bytes == 0
        return r, off // row_bytes

    # ---- forward push, layer by layer
    for k in range(world):
        fq, fx, _, _ = p2p_address_tables(plan, k, Lyr, nm, base_q, base_x, base_gq, base_gx)
        src = plan.parts[k]["send_idx"]
        for l in range(Lyr):
            assert fq[l].shape == src.shape
            for e in range(src.size):
                d, row = locate(int(fq[l][e]), base_q, 4 * H, Lyr * nm)
                d2, row2 = locate(int(fx[l][e]), base_x, 12, Lyr * nm)
                assert d == d2 and row == row2 and d != k
                assert l * nm + plan.parts[d]["n_own"] <= row < l * nm + nl[d]          # a halo row of layer l
                Q[d].reshape(-1, H)[row] = Qref[k][l, src[e]]
                X[d].reshape(-1, 3)[row] = Xref[k][l, src[e]]
    for k, p in enumerate(plan.parts):
        own, lid = plan.owner[p["halo"]], plan.local_id[p["halo"]]
        for l in range(Lyr):
            np.testing.assert_array_equal(Q[k][l, p["n_own"]:nl[k]], np.stack([Qref[o][l, i] for o, i in zip(own, lid)])
                                          if own.size else Q[k][l, :0])
            np.testing.assert_array_equal(Q[k][l, :p["n_own"]], Qref[k][l, :p["n_own"]])   # owned rows untouched
            np.testing.assert_array_equal(Q[k][l, nl[k]:], Qref[k][l, nl[k]:])             # padding untouched
            if own.size:
                np.testing.assert_array_equal(X[k][l, p["n_own"]:nl[k]], np.stack([Xref[o][l, i] for o, i in zip(own, lid)]))

    # ---- backward reduce push (both gx slots)
    gQ = [rng.standard_normal((nm, H)) for _ in range(world)]
    gX = [rng.standard_normal((2, nm, 3)) for _ in range(world)]
    outQ = [g.copy() for g in gQ]
    outX = [g.copy() for g in gX]
    for slot in range(2):
        for k, p in enumerate(plan.parts):
            _, _, bq, bx = p2p_address_tables(plan, k, Lyr, nm, base_q, base_x, base_gq, base_gx)
            assert bq.shape[0] == p["halo"].size
            for j in range(p["halo"].size):
                o, row = locate(int(bq[j]), base_gq, 4 * H, nm)
                o2, row2 = locate(int(bx[slot][j]), base_gx, 12, 2 * nm)
                assert o == o2 == plan.owner[p["halo"][j]] and row2 == slot * nm + row and row < plan.parts[o]["n_own"]
                if slot == 0:
                    outQ[o][row] += gQ[k][p["n_own"] + j]
                outX[o].reshape(-1, 3)[row2] += gX[k][slot, p["n_own"] + j]
    # reference: scatter-add of every user's halo rows by global node id
    for slot in range(2):
        accQ, accX = np.zeros((x.shape[0], H)), np.zeros((x.shape[0], 3))
        for k, p in enumerate(plan.parts):
            np.add.at(accQ, p["halo"], gQ[k][p["n_own"]:nl[k]])
            np.add.at(accX, p["halo"], gX[k][slot, p["n_own"]:nl[k]])
        for k, p in enumerate(plan.parts):
            if slot == 0:
                np.testing.assert_allclose(outQ[k][:p["n_own"]], gQ[k][:p["n_own"]] + accQ[p["owned"]], rtol=0, atol=1e-12)
            np.testing.assert_allclose(outX[k][slot, :p["n_own"]], gX[k][slot, :p["n_own"]] + accX[p["owned"]], rtol=0,
                                       atol=1e-12)


# ------------------------------------------------------------------------------------- device-built slab plan (host logic)
class _FakeGraph:
    """What CsrGraph.from_radius returns, from the numpy oracle (the CPU tests have no GPU to build it on)."""

    def __init__(self, cloud: torch.Tensor, r: float):
        from oracle import radius_graph_oracle as rgo
        n = cloud.size(0)
        o = rgo.radius_graph_csr(cloud.numpy(), np.array([0, n]), r, 0.0)
        self.N = self.Nl = n
        self.B, self.Fe = 1, 2
        self.E = int(o["row"].shape[0])
        self.row, self.col = torch.from_numpy(o["row"].astype(np.int32)), torch.from_numpy(o["col"].astype(np.int32))
        deg = np.bincount(o["row"], minlength=n)
        self.rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32))
        self.edge_attr = torch.from_numpy(np.stack([o["length"], o["length"]], 1))
        self.batch = torch.zeros(n, dtype=torch.int32)
        self.dinv = torch.from_numpy((1.0 / np.maximum(deg, 1)).astype(np.float32))
        self.inv_nb = torch.ones(1)


def _device_plan_worker(rank, world, port, x, r, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastegnn_b200.partitioned import DeviceSlabPlan
        plan = DeviceSlabPlan(torch.from_numpy(x), r, world, rank, build_graph=lambda c: _FakeGraph(c, r))
        g = plan.graph
        rows = plan.local_rows.numpy()
        out[rank] = dict(rows=rows, n_own=g.N, row=rows[g.row.numpy()], col=rows[g.col.numpy()],
                         length=g.edge_attr[:, 0].numpy(), dinv=g.dinv.numpy(), order=plan.order.numpy(),
                         parts=[dict(p) for p in plan.parts], owner=plan.owner, local_id=plan.local_id)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_device_slab_plan_reassembles_the_global_radius_graph(world):
    """DeviceSlabPlan: every rank builds its own slab's graph from (owned + candidate) points; together the slabs must
    hold exactly the global radius graph, each edge once at the owner of its row, with halo lists / send lists that
    mirror each other -- the same invariants test_slab_plan_reassembles_the_global_graph checks for the numpy plan."""
    from oracle import radius_graph_oracle as rgo
    rng = np.random.default_rng(4)
    n, r = 600, 0.2
    x = (rng.random((n, 3)) * np.array([3.0, 1.0, 1.0])).astype(np.float32)
    ref = rgo.radius_graph_csr(x, np.array([0, n]), r, 0.0)
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    if world == 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(_free_port())
        _device_plan_worker(0, 1, int(os.environ["MASTER_PORT"]), x, r, out)
    else:
        mp.spawn(_device_plan_worker, args=(world, _free_port(), x, r, out), nprocs=world, join=True)
    got = set()
    owned_all = []
    for k in range(world):
        o = out[k]
        owned_all.append(o["rows"][:o["n_own"]])
        assert np.isin(o["row"], o["rows"][:o["n_own"]]).all()                   # rows are owned
        pairs = set(zip(o["row"].tolist(), o["col"].tolist()))
        assert len(pairs) == o["row"].size and not (pairs & got)
        got |= pairs
        d = np.linalg.norm(x[o["row"]] - x[o["col"]], axis=1)
        np.testing.assert_allclose(o["length"], d, rtol=1e-5, atol=1e-7)
        deg = np.bincount(np.searchsorted(o["rows"][:o["n_own"]], o["row"], sorter=np.argsort(o["rows"][:o["n_own"]])),
                          minlength=o["n_own"])
        halo_rows = o["rows"][o["n_own"]:]
        assert np.array_equal(np.unique(o["col"][~np.isin(o["col"], o["rows"][:o["n_own"]])]), np.unique(halo_rows))
        parts, owner = o["parts"], o["owner"]
        assert np.array_equal(o["order"][parts[k]["halo"]], halo_rows)            # halo in (owner rank, owner-local id) order
        assert np.all(np.diff(parts[k]["halo"]) > 0) and (owner[parts[k]["halo"]] != k).all()
        assert parts[k]["recv_counts"].sum() == parts[k]["halo"].size
    assert got == set(zip(ref["row"].tolist(), ref["col"].tolist()))
    assert sorted(np.concatenate(owned_all).tolist()) == list(range(n))
    for k in range(world):                                                        # send lists mirror the receivers' halos
        pk = out[k]["parts"][k]
        for d in range(world):
            o0 = int(pk["send_counts"][:d].sum())
            sent = out[k]["order"][out[k]["parts"][k]["send_idx"][o0:o0 + int(pk["send_counts"][d])] + int(n * k // world)]
            hd = out[d]["parts"][d]["halo"]
            want = out[d]["order"][hd[out[d]["owner"][hd] == k]]
            assert np.array_equal(sent, want), (k, d)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_fused_halo_tables_in_a_simulated_address_space(world):
    """Tables of the fused (payload + signal) halo kernels: forward stores land in the users' halo rows; the reverse halo
    stores every user's halo-gradient row into a distinct slot of the owner's receive buffer, and the owner's (rows, ptr,
    slots) lists add exactly those slots -- in a FIXED order -- to the right owned rows."""
    from fastegnn_b200.partitioned import SlabPlan, fused_halo_tables
    x, ei = _cloud(n=500, deg=10, seed=3)
    plan = SlabPlan(x, ei, world)
    Lyr, H = 2, 64
    nl = [p["n_own"] + p["halo"].size for p in plan.parts]
    nm = max(nl)
    cap = max(1, max(int(p["send_counts"].sum()) for p in plan.parts))
    mk = lambda j: [((r + 1) << 40) + (j << 36) for r in range(world)]
    base_q, base_x, base_rq, base_rx = mk(0), mk(1), mk(2), mk(3)
    rng = np.random.default_rng(2)
    gQ = [rng.standard_normal((nm, H)) for _ in range(world)]
    gX = [rng.standard_normal((nm, 3)) for _ in range(world)]
    for par in range(2):
        RQ = [np.full((2, cap, H), np.nan) for _ in range(world)]
        RX = [np.full((2, cap, 3), np.nan) for _ in range(world)]
        tabs = [fused_halo_tables(plan, k, Lyr, nm, cap, base_q, base_x, base_rq, base_rx) for k in range(world)]
        for k, p in enumerate(plan.parts):                                        # every user pushes its halo rows
            _, _, bq, bx, _, _, _ = tabs[k]
            for j in range(p["halo"].size):
                o = int(bq[par][j] >> 40) - 1
                off = int(bq[par][j]) - base_rq[o]
                assert off % (4 * H) == 0 and o == plan.owner[p["halo"][j]]
                slot = off // (4 * H)
                assert par * cap <= slot < (par + 1) * cap
                assert int(bx[par][j]) - base_rx[o] == slot * 12
                assert np.isnan(RQ[o].reshape(-1, H)[slot]).all()                  # no two rows share a slot
                RQ[o].reshape(-1, H)[slot] = gQ[k][p["n_own"] + j]
                RX[o].reshape(-1, 3)[slot] = gX[k][p["n_own"] + j]
        accQ, accX = np.zeros((x.shape[0], H)), np.zeros((x.shape[0], 3))
        for k, p in enumerate(plan.parts):
            np.add.at(accQ, p["halo"], gQ[k][p["n_own"]:nl[k]])
            np.add.at(accX, p["halo"], gX[k][p["n_own"]:nl[k]])
        for k, p in enumerate(plan.parts):                                        # every owner applies its slots
            _, _, _, _, rows, ptr, slots = tabs[k]
            outQ, outX = gQ[k].copy(), gX[k].copy()
            assert ptr[0] == 0 and ptr[-1] == slots.size == int(p["send_counts"].sum())
            for i, rrow in enumerate(rows):
                sl = slots[ptr[i]:ptr[i + 1]]
                assert np.all(np.diff(sl) > 0)                                     # (user rank, halo order): fixed order
                for s_ in sl:
                    outQ[rrow] += RQ[k][par, s_]
                    outX[rrow] += RX[k][par, s_]
            np.testing.assert_allclose(outQ[:p["n_own"]], gQ[k][:p["n_own"]] + accQ[p["owned"]], rtol=0, atol=1e-12)
            np.testing.assert_allclose(outX[:p["n_own"]], gX[k][:p["n_own"]] + accX[p["owned"]], rtol=0, atol=1e-12)
