"""Generate tests/golden/graph_*.npz: outputs of the UNMODIFIED reference's edge selection code.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_graph.py

The dataset modules cannot be imported (MDAnalysis, torch_geometric, torch_cluster are absent), so the two
``cutoff_edge`` methods are extracted with ``ast`` and executed verbatim:
  * datasets/simulation/dataset.py:96-101  cutoff_edge(self, edge_index, loc_0)   (also datasets/protein/dataset.py:208-213)
  * datasets/nbody/dataset.py:102-113      cutoff_edge(self, loc_0)
``radius_graph`` itself (torch_cluster, third party, absent) is replaced by the candidate list of
oracle/radius_graph_oracle.py in torch_cluster's emission order -- see that module's header: that part is unpinned.

Test infrastructure; never imported by the product.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle import radius_graph_oracle as rgo  # noqa: E402


def _extract_method(path: str, cls: str, name: str):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    fn.decorator_list = []
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ns = {"torch": torch}
                    exec(compile(mod, path, "exec"), ns)
                    return ns[name], (fn.lineno, fn.end_lineno)
    raise SystemExit(f"{cls}.{name} not found in {path}")


def _class_of(path: str, method: str) -> str:
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and any(isinstance(f, ast.FunctionDef) and f.name == method for f in node.body):
            return node.name
    raise SystemExit(f"no class with {method} in {path}")


def main():
    os.makedirs(OUT, exist_ok=True)
    sim_path = os.path.join(REF, "datasets/simulation/dataset.py")
    nb_path = os.path.join(REF, "datasets/nbody/dataset.py")
    sim_cut, sim_lines = _extract_method(sim_path, _class_of(sim_path, "cutoff_edge"), "cutoff_edge")
    nb_cut, nb_lines = _extract_method(nb_path, _class_of(nb_path, "cutoff_edge"), "cutoff_edge")

    # ---- Water-3D style: radius graph (oracle candidates, torch_cluster order) -> reference cutoff_edge
    rng = np.random.default_rng(7)
    for tag, n, r, cr in (("sim_cr0", 400, 0.035, 0.0), ("sim_cr25", 400, 0.035, 0.25), ("sim_cr50", 300, 0.05, 0.5)):
        rho = 20.0 / (4.0 / 3.0 * np.pi * r ** 3)
        side = (n / rho) ** (1.0 / 3.0)
        x = (rng.random((n, 3)) * side).astype(np.float32)
        ptr = np.array([0, n])
        ei = rgo.reference_order_edge_list(x, ptr, r)
        me = types.SimpleNamespace(cutoff_rate=cr)
        out = sim_cut(me, torch.from_numpy(ei), torch.from_numpy(x))
        np.savez_compressed(os.path.join(OUT, f"graph_{tag}.npz"), x=x, ptr=ptr, r=np.float64(r), cutoff_rate=np.float64(cr),
                            ref_edge_index=out.numpy().astype(np.int64),
                            source=np.array(f"datasets/simulation/dataset.py:{sim_lines[0]}-{sim_lines[1]}"))
        print(tag, "candidates", ei.shape[1], "kept", out.shape[1])

    # ---- N-body style: complete graph, topk shortest (one graph per call, as the dataset does)
    for tag, n, cr in (("nbody_n5_cr50", 5, 0.5), ("nbody_n20_cr50", 20, 0.5), ("nbody_n20_cr0", 20, 0.0)):
        sigma = (n / 5.0) ** (1.0 / 3.0) + 0.1
        x = (rng.standard_normal((n, 3)) * sigma).astype(np.float32)
        me = types.SimpleNamespace(cutoff_rate=cr)
        out = nb_cut(me, torch.from_numpy(x))
        np.savez_compressed(os.path.join(OUT, f"graph_{tag}.npz"), x=x, ptr=np.array([0, n]), r=np.float64(np.inf),
                            cutoff_rate=np.float64(cr), ref_edge_index=out.numpy().astype(np.int64),
                            source=np.array(f"datasets/nbody/dataset.py:{nb_lines[0]}-{nb_lines[1]}"))
        print(tag, "kept", out.shape[1])


if __name__ == "__main__":
    main()
