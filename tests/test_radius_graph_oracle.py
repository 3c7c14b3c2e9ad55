"""CPU: oracle/radius_graph_oracle.py against the outputs of the reference's own edge-selection code
(tests/golden/graph_*.npz, made by oracle/make_golden_graph.py from datasets/*/dataset.py cutoff_edge)."""
import glob
import os

import numpy as np
import pytest

from oracle import radius_graph_oracle as rgo

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[len("graph_"):-4] for p in glob.glob(os.path.join(GOLD, "graph_*.npz")))


def load(tag):
    z = np.load(os.path.join(GOLD, f"graph_{tag}.npz"))
    return z["x"], z["ptr"], float(z["r"]), float(z["cutoff_rate"]), z["ref_edge_index"]


def test_fixtures_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("tag", [c for c in CASES if c.startswith("sim")])
def test_csr_equals_stable_row_sort_of_reference_cutoff_edge(tag):
    """graph_prep(reference list) == the oracle's CSR, index for index: same edge set, same tie choice at the cut,
    same order inside every row (ascending length, ties by col)."""
    x, ptr, r, cr, ref = load(tag)
    g = rgo.radius_graph_csr(x, ptr, r, cr)
    order = np.argsort(ref[0], kind="stable")                  # models/FastEGNN.py scatter-by-row == stable sort by row
    assert g["row"].shape[0] == ref.shape[1] == int(g["n_candidates"] * (1 - cr))
    np.testing.assert_array_equal(g["row"], ref[0][order].astype(np.int32))
    np.testing.assert_array_equal(g["col"], ref[1][order].astype(np.int32))
    # lengths are what torch.norm gives, to fp32 rounding
    d = x[ref[0][order]].astype(np.float64) - x[ref[1][order]].astype(np.float64)
    np.testing.assert_allclose(g["length"], np.sqrt((d * d).sum(1)), rtol=3e-7)


@pytest.mark.parametrize("tag", [c for c in CASES if c.startswith("nbody")])
def test_complete_graph_topk_selects_the_same_pairs(tag):
    """datasets/nbody/dataset.py:102-113 (cdist + topk): the same SET of ordered pairs (topk's order among equal
    lengths is unspecified, and the cut never splits a twin pair at these sizes)."""
    x, ptr, r, cr, ref = load(tag)
    g = rgo.radius_graph_csr(x, ptr, np.inf, cr)
    mine = set(zip(g["row"].tolist(), g["col"].tolist()))
    theirs = set(zip(ref[0].tolist(), ref[1].tolist()))
    assert mine == theirs
    assert np.all(np.diff(g["row"]) >= 0)
    for i in range(x.shape[0]):                                # inside a row: ascending (length, col)
        s, e = g["rowptr"][i], g["rowptr"][i + 1]
        keys = list(zip(g["length"][s:e].tolist(), g["col"][s:e].tolist()))
        assert keys == sorted(keys)


def test_multi_graph_batches_never_connect_across_graphs():
    rng = np.random.default_rng(0)
    sizes = [7, 1, 0, 30, 12]
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    x = rng.random((ptr[-1], 3)).astype(np.float32) * 0.5
    g = rgo.radius_graph_csr(x, ptr, 0.2, 0.3)
    gid = np.searchsorted(ptr[1:], np.arange(ptr[-1]), side="right")
    assert np.all(gid[g["row"]] == gid[g["col"]]) and np.all(g["row"] != g["col"])
    full = rgo.radius_graph_csr(x, ptr, 0.2, 0.0)
    for b in range(len(sizes)):
        eb = int(np.sum(gid[full["row"]] == b))
        assert int(np.sum(gid[g["row"]] == b)) == int(eb * (1 - 0.3))
