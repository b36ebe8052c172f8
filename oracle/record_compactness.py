#!/usr/bin/env python
"""Record the reference's compactness updaters as a fixture (tests/golden/compactness_*.npz).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/record_compactness.py

For every state of a ReCom chain of the UNMODIFIED reference (so the values have gone
through its incremental flow updaters, updaters/compactness.py) it stores the plan and what
the reference reports for ``exterior_boundaries``, ``interior_boundaries``, ``perimeter``,
``area`` and ``polsby_popper`` per district, plus ``boundary_nodes``:

* ``grid12``: the reference's own ``Grid((12, 12))`` (integer geometry, grid.py:142-181)
* ``delaunay300``: a triangulation with float ``shared_perim`` / ``boundary_perim`` / ``area``
"""

from __future__ import annotations

import json
import os
import random
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from record_reference import import_reference  # noqa: E402


def float_geometry_graph(n=300, seed=11):
    """Delaunay graph with made-up float geometry: edge lengths as shared_perim, hull nodes
    as boundary nodes."""
    import networkx as nx
    from scipy.spatial import ConvexHull, Delaunay

    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    s = Delaunay(pts).simplices
    e = np.unique(np.sort(np.concatenate([s[:, [0, 1]], s[:, [1, 2]], s[:, [0, 2]]]), axis=1), axis=0)
    g = nx.Graph()
    hull = set(ConvexHull(pts).vertices.tolist())
    for i in range(n):
        g.add_node(i, population=int(rng.integers(50, 150)), area=float(rng.random() + 0.5), boundary_node=i in hull)
        if i in hull:
            g.nodes[i]["boundary_perim"] = float(rng.random() + 0.1)
    for u, v in e.tolist():
        g.add_edge(u, v, shared_perim=float(np.hypot(*(pts[u] - pts[v]))))
    return g


def record(name, steps=25):
    import_reference()
    from gerrychain import Graph, GeographicPartition, MarkovChain
    from gerrychain.accept import always_accept
    from gerrychain.constraints import within_percent_of_ideal_population
    from gerrychain.grid import Grid
    from gerrychain.metrics import polsby_popper
    from gerrychain.proposals import recom
    from gerrychain.tree import recursive_tree_part
    from gerrychain.updaters import Tally

    random.seed(20240)
    if name == "grid12":
        init = Grid((12, 12))
        nxg = init.graph
        pop_col, eps = "population", 0.2
    else:
        nxg = float_geometry_graph()
        graph = Graph.from_networkx(nxg)
        k = 5
        total = sum(nxg.nodes[v]["population"] for v in nxg.nodes)
        plan = recursive_tree_part(graph, range(k), total / k, "population", 0.1)
        init = GeographicPartition(graph, plan, {"population": Tally("population"), "polsby_popper": polsby_popper})
        pop_col, eps = "population", 0.1
    k = len(init.parts)
    total = sum(init["population"].values())
    chain = MarkovChain(partial(recom, pop_col=pop_col, pop_target=total / k, epsilon=eps, node_repeats=1),
                        [within_percent_of_ideal_population(init, eps)], always_accept, init, steps)
    labels = list(init.graph.nodes)
    parts = sorted(init.parts)
    rows = {key: [] for key in ("plan", "exterior", "interior", "perimeter", "area", "polsby")}
    for state in chain:
        rows["plan"].append([parts.index(state.assignment.mapping[v]) for v in labels])
        ext = state["exterior_boundaries"]
        rows["exterior"].append([float(ext[p]) for p in parts])  # defaultdict: 0 for parts without boundary nodes
        rows["interior"].append([float(state["interior_boundaries"][p]) for p in parts])
        rows["perimeter"].append([float(state["perimeter"][p]) for p in parts])
        rows["area"].append([float(state["area"][p]) for p in parts])
        rows["polsby"].append([float(state["polsby_popper"][p]) for p in parts])
    bn = sorted(labels.index(v) for v in init["boundary_nodes"])
    node_attrs = [dict(init.graph.nodes[v]) for v in labels]
    edges = [(labels.index(u), labels.index(v), float(d["shared_perim"]), isinstance(d["shared_perim"], int))
             for u, v, d in init.graph.edges(data=True)]
    out = os.path.join(REPO, "tests", "golden", f"compactness_{name}.npz")
    np.savez_compressed(
        out,
        meta=json.dumps(dict(name=name, parts=[int(p) for p in parts], steps=steps,
                             labels=[list(v) if isinstance(v, tuple) else v for v in labels])),
        plan=np.array(rows["plan"], dtype=np.uint8),
        exterior=np.array(rows["exterior"]), interior=np.array(rows["interior"]),
        perimeter=np.array(rows["perimeter"]), area=np.array(rows["area"]), polsby=np.array(rows["polsby"]),
        boundary_nodes=np.array(bn, dtype=np.int64),
        node_boundary=np.array([bool(a.get("boundary_node", False)) for a in node_attrs]),
        node_boundary_perim=np.array([float(a.get("boundary_perim", 0.0)) for a in node_attrs]),
        node_area=np.array([float(a["area"]) for a in node_attrs]),
        node_population=np.array([int(a["population"]) for a in node_attrs], dtype=np.int64),
        edge_uv=np.array([(u, v) for u, v, _, _ in edges], dtype=np.int64),
        edge_shared_perim=np.array([w for _, _, w, _ in edges]),
        integral=np.array(all(i for _, _, _, i in edges)),
    )
    print("wrote", out, len(rows["plan"]), "states")


if __name__ == "__main__":
    for nm in ("grid12", "delaunay300"):
        record(nm)
