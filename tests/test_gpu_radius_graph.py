"""GPU: on-device graph construction (fegnn_radius_graph_count / _fill through CsrGraph.from_radius) against
oracle/radius_graph_oracle.py and the reference's own cutoff_edge outputs.  Index work: everything bit-exact."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import radius_graph_oracle as rgo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def build(x, ptr, r, cr, Fe=2):
    from fastegnn_b200 import CsrGraph
    ptr = np.asarray(ptr)
    B = len(ptr) - 1
    batch = torch.from_numpy(np.repeat(np.arange(B), np.diff(ptr))).long().to(DEV)
    g = CsrGraph.from_radius(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(DEV), batch, B, r, cr, Fe)
    torch.cuda.synchronize()
    return g


def check(x, ptr, r, cr, Fe=2):
    g = build(x, ptr, r, cr, Fe)
    o = rgo.radius_graph_csr(x, ptr, r, cr)
    assert g.n_candidates == o["n_candidates"]
    assert g.E == o["row"].shape[0]
    np.testing.assert_array_equal(g.rowptr.cpu().numpy(), o["rowptr"])
    np.testing.assert_array_equal(g.row.cpu().numpy(), o["row"])
    np.testing.assert_array_equal(g.col.cpu().numpy(), o["col"])
    if Fe:
        ea = g.edge_attr.cpu().numpy()
        assert ea.shape == (g.E, Fe)
        for f in range(Fe):
            np.testing.assert_array_equal(ea[:, f].view(np.uint32), o["length"].view(np.uint32))     # bit for bit
    deg = np.diff(o["rowptr"])
    np.testing.assert_array_equal(g.dinv.cpu().numpy(), (1.0 / np.maximum(deg, 1)).astype(np.float32))
    n = np.diff(np.asarray(ptr))
    np.testing.assert_array_equal(g.inv_nb.cpu().numpy(), (1.0 / np.maximum(n, 1)).astype(np.float32))
    np.testing.assert_array_equal(g.gptr.cpu().numpy(), np.asarray(ptr, dtype=np.int32))
    return g, o


@pytest.mark.parametrize("tag", sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLD, "graph_*.npz"))))
def test_golden_fixtures_of_the_reference_cutoff_edge(tag):
    z = np.load(os.path.join(GOLD, f"graph_{tag}.npz"))
    x, ptr, r, cr, ref = z["x"], z["ptr"], float(z["r"]), float(z["cutoff_rate"]), z["ref_edge_index"]
    g, _ = check(x, ptr, r, cr)
    if tag.startswith("sim"):          # == graph_prep (stable sort by row) of the reference's own edge list
        order = np.argsort(ref[0], kind="stable")
        np.testing.assert_array_equal(g.row.cpu().numpy(), ref[0][order].astype(np.int32))
        np.testing.assert_array_equal(g.col.cpu().numpy(), ref[1][order].astype(np.int32))
    else:
        assert set(zip(g.row.tolist(), g.col.tolist())) == set(zip(ref[0].tolist(), ref[1].tolist()))


@pytest.mark.parametrize("cr", [0.0, 0.25, 0.5])
def test_ragged_batch_with_empty_and_single_node_graphs(cr):
    rng = np.random.default_rng(3)
    sizes = [40, 1, 0, 257, 2, 0, 90, 33]
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    x = (rng.random((ptr[-1], 3)) * np.array([1.0, 0.4, 2.5])).astype(np.float32)
    x[ptr[3]:ptr[4]] += 50.0                     # graphs far apart and overlapping in space: grids are per graph
    check(x, ptr, 0.3, cr)


@pytest.mark.parametrize("cr", [0.5, 0.25, 0.0])
def test_nbody_complete_graphs_topk(cr):
    """datasets/nbody/dataset.py:102-113 at config-2 shape: 100 graphs x 100 particles, r = inf (degree 99: the
    heap-sort path of the row ordering)."""
    rng = np.random.default_rng(11)
    n, B = 100, 100
    x = (rng.standard_normal((n * B, 3)) * 2.8).astype(np.float32)
    ptr = np.arange(B + 1) * n
    g, _ = check(x, ptr, math.inf, cr)
    assert g.E == B * int(n * (n - 1) * (1 - cr))


def test_lattice_all_lengths_tied_and_strict_radius():
    """Points on a grid with spacing exactly r: d2 == r*r pairs are NOT edges (strict <); with a larger radius every
    length occurs many times and the cut goes through a run of ties -> (length, col, row) decides."""
    k = 9
    ax = np.arange(k, dtype=np.float32) * np.float32(0.25)
    x = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ptr = np.array([0, x.shape[0]])
    g, _ = check(x, ptr, 0.25, 0.0)
    assert g.E == 0
    check(x, ptr, 0.3, 0.0)
    check(x, ptr, 0.3, 0.37)
    check(x, ptr, 0.51, 0.61)
    check(np.zeros((50, 3), dtype=np.float32), np.array([0, 50]), 1.0, 0.3)     # coincident points: all lengths 0


def test_sparse_cloud_in_a_huge_box_clamps_the_cell_grid():
    rng = np.random.default_rng(5)
    x = (rng.random((3000, 3)) * np.array([1000.0, 10.0, 0.0])).astype(np.float32)        # > 255 cells along x, flat in z
    check(x, np.array([0, 3000]), 0.7, 0.1)
    x2 = np.concatenate([rng.random((500, 3)), rng.random((500, 3)) + 1e4]).astype(np.float32)   # two distant clusters
    check(x2, np.array([0, 1000]), 0.2, 0.0)


def test_negative_coordinates_and_dense_rows():
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((1500, 3)) * 0.2 - 3.0).astype(np.float32)                  # mean degree in the hundreds
    check(x, np.array([0, 700, 1500]), 0.25, 0.2, Fe=1)


def test_water3d_full_size_and_model_consumes_the_prebuilt_graph():
    """Config-4 shape (8 000 particles, ~1.8e5 directed edges): bit-exact CSR, then FastEGNN.forward with the prebuilt
    graph equals forward on the int64 edge list exported from it."""
    from bench import make_cloud
    from fastegnn_b200 import FastEGNN
    data = make_cloud(8000, 25.0, 3, 0, [0, -1, 0])
    x = data["loc_0"].numpy()
    g, o = check(x, np.array([0, 8000]), 0.035, 0.0)
    ref_pairs = set(zip(data["edge_index"][0].tolist(), data["edge_index"][1].tolist()))    # scipy cKDTree graph of bench.py
    assert abs(len(ref_pairs) - g.E) <= 4        # fp64 vs fp32 decision exactly at the radius
    torch.manual_seed(0)
    m = FastEGNN(node_feat_nf=2, node_attr_nf=0, edge_attr_nf=2, hidden_nf=64, virtual_channels=3, device=DEV,
                 gravity=[0, -1, 0])
    t = {k: data[k].to(DEV) for k in ("node_feat", "loc_0", "vel_0", "batch", "loc_mean")}
    from fastegnn_b200 import _lib
    _lib.set_precision("fp32")       # same CSR arrays either way; fp32 kernels so that only atomics order differs
    with torch.no_grad():
        x1, Z1 = m(node_feat=t["node_feat"], node_loc=t["loc_0"], node_vel=t["vel_0"], edge_index=g,
                   data_batch=t["batch"], loc_mean=t["loc_mean"], edge_attr=None)
        x2, Z2 = m(node_feat=t["node_feat"], node_loc=t["loc_0"], node_vel=t["vel_0"], edge_index=g.edge_index(),
                   data_batch=t["batch"], loc_mean=t["loc_mean"], edge_attr=g.edge_attr)
    _lib.set_precision("tf32")
    assert torch.isfinite(x1).all()
    assert (x1 - x2).abs().max().item() <= 1e-5 * x2.abs().max().item()
    assert (Z1 - Z2).abs().max().item() <= 1e-5 * Z2.abs().max().item()
    with pytest.raises(TypeError):
        m(node_feat=t["node_feat"], node_loc=t["loc_0"], node_vel=t["vel_0"], edge_index=g, data_batch=t["batch"],
          loc_mean=t["loc_mean"], edge_attr=g.edge_attr)


def test_empty_input():
    from fastegnn_b200 import CsrGraph
    g = CsrGraph.from_radius(torch.zeros(0, 3, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV), 0, 1.0)
    assert g.E == 0 and g.N == 0
    g = build(np.zeros((1, 3), dtype=np.float32), np.array([0, 1]), 1.0, 0.0)
    assert g.E == 0 and g.rowptr.tolist() == [0, 0]
