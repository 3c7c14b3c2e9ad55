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
