#!/usr/bin/env python
"""Seed plans for the Delaunay configurations, from the reference's recursive_tree_part
(/root/reference/gerrychain/tree.py:971-1085) with random.seed(0).

TEST INFRASTRUCTURE ONLY; run in the build container:  python oracle/make_seed_plans.py
Writes tests/golden/plan_delaunay<N>_k<k>.npy (uint8 district index per node).
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from record_reference import import_reference  # noqa: E402
from gerrychain_b200 import synthetic as syn  # noqa: E402


def make(n_nodes, k, eps, pop_range=(500, 2500)):
    import_reference()
    from gerrychain import Graph
    from gerrychain.tree import recursive_tree_part

    g = syn.delaunay_graph(n_nodes, seed=0, pop_range=pop_range)
    tot = sum(g.nodes[v]["population"] for v in g)
    random.seed(0)
    plan = recursive_tree_part(Graph.from_networkx(g), range(k), tot / k, "population", eps)
    vec = np.array([plan[v] for v in g.nodes], dtype=np.uint8)
    out = os.path.join(REPO, "tests", "golden", f"plan_delaunay{n_nodes}_k{k}.npy")
    np.save(out, vec)
    print("wrote", out, np.bincount(vec))


if __name__ == "__main__":
    if "--only-big" not in sys.argv:
        make(9000, 18, 0.02)
    if "--big" in sys.argv or "--only-big" in sys.argv:
        make(200000, 50, 0.05, pop_range=(20, 200))
