"""CPU restatement of the graph construction in front of the FastEGNN layer path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline) may import this module; the product
(fastegnn_b200/) builds graphs with the CUDA kernels of csrc/radius_graph.cu and never comes here.

What it follows in the reference
    datasets/simulation/dataset.py:80       edge_index = radius_graph(loc_0, r=0.035, max_num_neighbors=100000)
    datasets/simulation/dataset.py:96-101   cutoff_edge: torch.sort(edge_dist), keep int(E * (1 - cutoff_rate)) shortest
    datasets/simulation/dataset.py:82       edge_attr = torch.norm(loc_0[row] - loc_0[col], p=2, dim=1)
    datasets/nbody/dataset.py:102-113       complete graph (diagonal pushed to 1e18), topk of the n(n-1)(1-cr) shortest
    datasets/protein/dataset.py:146-156     contact graph (cutoff 10), self loops removed, same cutoff_edge (:208-213)
    utils/train.py:41-43                    edge_attr = cat[edge_attr, |loc_0[row] - loc_0[col]|]  (both columns = length)
    models/FastEGNN.py:279-294              scatter by row == CSR by row after a stable sort (graph_prep)

Third-party arithmetic that is NOT under /root/reference: torch_cluster==1.6.3 (requirements.txt) `radius_graph`.
Its published behaviour, restated here: for every target node i the sources j != i (loop=False) of the same example
with squared distance < r*r, at most max_num_neighbors of them (100000: never binding); the CUDA kernel scans j in
ascending index and emits the list grouped by target, edge_index = [source; target].  Because the graph is symmetric the
set of ordered pairs does not depend on which end is called "row".  **Parity unpinned for radius_graph itself**: the
package is absent from this image and the reference holds no fixture of its output; what IS pinned is `cutoff_edge`,
executed verbatim from the reference source on this oracle's candidate list (tests/test_radius_graph_oracle.py).

Definition used by both this oracle and the CUDA kernels (bit-exact between the two):
    d2(i,j)  = (dx*dx + dy*dy) + dz*dz, each operation rounded to fp32 (no fused multiply-add); length = sqrt(d2) in fp32
    edges    = {(i,j): i != j, same graph, d2 < fp32(r)*fp32(r)}
    keep     = per graph the int(E_b * (1 - cutoff_rate)) first in (length, col, row) order  [a stable sort of the
               target-grouped list by length]
    layout   = CSR by row; inside a row ascending (length, col)
"""
from __future__ import annotations

import math

import numpy as np


def _d2(x: np.ndarray, i: np.ndarray, j: np.ndarray) -> np.ndarray:
    d = x[i] - x[j]                                     # fp32 subtraction
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def candidate_pairs(x: np.ndarray, ptr: np.ndarray, r: float):
    """All ordered pairs (row i, col j) of the definition above, unordered; x fp32 [N,3], ptr int [B+1]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    r32 = np.float32(r)
    r2 = r32 * r32 if math.isfinite(r) else np.float32(np.inf)
    rows, cols = [], []
    for b in range(len(ptr) - 1):
        s, e = int(ptr[b]), int(ptr[b + 1])
        n = e - s
        if n < 2:
            continue
        if n <= 2048 or not math.isfinite(r):
            i, j = np.meshgrid(np.arange(s, e), np.arange(s, e), indexing="ij")
            i, j = i.reshape(-1), j.reshape(-1)
            m = i != j
            i, j = i[m], j[m]
        else:                                           # prune with a KD-tree (slack radius), decide in fp32 below
            from scipy.spatial import cKDTree
            pr = cKDTree(x[s:e].astype(np.float64)).query_pairs(float(r) * (1.0 + 1e-4) + 1e-30, output_type="ndarray")
            i = np.concatenate([pr[:, 0], pr[:, 1]]) + s
            j = np.concatenate([pr[:, 1], pr[:, 0]]) + s
        ok = _d2(x, i, j) < r2
        rows.append(i[ok])
        cols.append(j[ok])
    if not rows:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    return np.concatenate(rows).astype(np.int64), np.concatenate(cols).astype(np.int64)


def radius_graph_csr(x: np.ndarray, ptr: np.ndarray, r: float, cutoff_rate: float = 0.0):
    """-> dict(rowptr int32 [N+1], row int32 [E], col int32 [E], length fp32 [E], n_candidates)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    N = x.shape[0]
    row, col = candidate_pairs(x, ptr, r)
    length = np.sqrt(_d2(x, row, col)).astype(np.float32)
    n_cand = int(row.size)
    if cutoff_rate > 0.0 and n_cand:
        graph_of = np.searchsorted(np.asarray(ptr)[1:], row, side="right")
        keep = np.zeros(n_cand, dtype=bool)
        for b in range(len(ptr) - 1):
            idx = np.nonzero(graph_of == b)[0]
            k = int(idx.size * (1 - cutoff_rate))                      # datasets/simulation/dataset.py:99
            order = np.lexsort((row[idx], col[idx], length[idx]))      # (length, col, row)
            keep[idx[order[:k]]] = True
        row, col, length = row[keep], col[keep], length[keep]
    order = np.lexsort((col, length, row))                             # CSR: row, then (length, col)
    row, col, length = row[order], col[order], length[order]
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    rowptr = np.cumsum(rowptr)
    return dict(rowptr=rowptr.astype(np.int32), row=row.astype(np.int32), col=col.astype(np.int32), length=length,
                n_candidates=n_cand)


def reference_order_edge_list(x: np.ndarray, ptr: np.ndarray, r: float):
    """The candidate list in the order torch_cluster's CUDA radius kernel emits it (grouped by target = col ascending,
    sources = row ascending): edge_index int64 [2,E] = [source; target]."""
    row, col = candidate_pairs(x, ptr, r)
    order = np.lexsort((row, col))
    return np.stack([row[order], col[order]])
