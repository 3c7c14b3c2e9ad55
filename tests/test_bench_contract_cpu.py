"""CPU: the bench.py contract that does not need a GPU -- the reference arm's JSON line and the synthetic workloads."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nodes", "600",
                        "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == dict(value=d["value"], unit="edges/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-6 * d["value"]


@pytest.mark.parametrize("name,nodes", [("water3d", 700), ("water3d_b20", 300), ("nbody5", 0), ("nbody100", 0), ("protein", 0), ("large", 3000)])
def test_workloads_have_reference_shaped_batches(name, nodes):
    """What utils/train.py:32-53 hands to the model: int64 edge_index without self loops, non-decreasing int64 batch whose
    last entry is B-1, loc_mean [B,3,C], edge_attr [E,2] = (length, length) (datasets + utils/train.py:41-43)."""
    sys.path.insert(0, ROOT)
    import bench
    data, hp = bench.make_workload(name, seed=0, nodes=nodes)
    N, E, B, C = data["loc_0"].size(0), data["edge_index"].size(1), data["n_graphs"], data["C"]
    assert data["edge_index"].dtype == torch.int64 and data["batch"].dtype == torch.int64
    assert data["node_feat"].shape == (N, 2) and data["vel_0"].shape == (N, 3) and data["loc_t"].shape == (N, 3)
    assert data["edge_attr"].shape == (E, 2) and data["loc_mean"].shape == (B, 3, C)
    assert int(data["batch"][-1]) == B - 1 and bool((data["batch"][1:] >= data["batch"][:-1]).all())
    assert sum(data["sizes"]) == N and len(data["sizes"]) == B
    r, c = data["edge_index"]
    assert bool((r != c).all()) and int(r.max()) < N and int(c.max()) < N
    assert bool((data["batch"][r] == data["batch"][c]).all())                      # no edge crosses graphs
    length = (data["loc_0"][r] - data["loc_0"][c]).norm(dim=1)
    np.testing.assert_allclose(data["edge_attr"][:, 0].numpy(), length.numpy(), rtol=1e-5, atol=1e-7)
    assert torch.equal(data["edge_attr"][:, 0], data["edge_attr"][:, 1])
    assert set(hp) == {"sigma", "weight", "sample"}
    if name == "nbody100":
        assert E == 100 * int(100 * 99 * 0.5)                                      # datasets/nbody/dataset.py:107


def test_radius_oracle_kdtree_path_equals_brute_force():
    """oracle/radius_graph_oracle.py switches to KD-tree pruning above 2 048 nodes per graph; both paths must give the
    same CSR (the decision itself is always the fp32 d2 < r^2 test)."""
    from oracle import radius_graph_oracle as rgo
    rng = np.random.default_rng(2)
    x = rng.random((2300, 3)).astype(np.float32)
    ptr = np.array([0, 2300])
    a = rgo.radius_graph_csr(x, ptr, 0.08, 0.3)
    # brute force through two graphs-worth of slicing is not possible (one graph); emulate by splitting the pair list
    i, j = np.meshgrid(np.arange(2300), np.arange(2300), indexing="ij")
    i, j = i.reshape(-1), j.reshape(-1)
    m = i != j
    i, j = i[m], j[m]
    d = x[i] - x[j]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    ok = d2 < np.float32(0.08) * np.float32(0.08)
    assert int(ok.sum()) == a["n_candidates"]
    i, j, ln = i[ok], j[ok], np.sqrt(d2[ok]).astype(np.float32)
    keep = np.lexsort((i, j, ln))[:int(ok.sum() * (1 - 0.3))]
    i, j, ln = i[keep], j[keep], ln[keep]
    o = np.lexsort((j, ln, i))
    np.testing.assert_array_equal(a["row"], i[o].astype(np.int32))
    np.testing.assert_array_equal(a["col"], j[o].astype(np.int32))
    np.testing.assert_array_equal(a["length"].view(np.uint32), ln[o].view(np.uint32))
