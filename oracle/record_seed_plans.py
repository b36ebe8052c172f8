#!/usr/bin/env python
"""Record the reference's recursive_tree_part (seed plans) as replay fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/record_seed_plans.py

Same recipe as oracle/record_reference.py: the UNMODIFIED reference
(/root/reference/gerrychain/tree.py:971-1085) is run with wrapped injection points
(spanning_tree_fn -> per-edge weights, choice -> BFS root, cut_choice -> chosen cut), one
record per split, plus the plan it returns.  Writes tests/golden/seedplan_*.npz.
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
from oracle.record_reference import import_reference  # noqa: E402


def record(nx_graph, k, eps, pop_col, path, seed):
    import_reference()
    from gerrychain import Graph
    from gerrychain.tree import (bipartition_tree, random_spanning_tree, recursive_tree_part,
                                 _region_preferred_max_weight_choice)
    from gerrychain_b200.csr import csr_from_networkx

    graph = Graph.from_networkx(nx_graph)
    csr = csr_from_networkx(graph, pop_col)
    index = csr.node_index()
    eid = csr.edge_id_lookup()

    def edge_id(e):
        return eid[tuple(sorted((index[e[0]], index[e[1]])))]

    log = []

    def rec_mst(g, region_surcharge=None):
        t = random_spanning_tree(g, region_surcharge=region_surcharge)
        log.append(("tree", {e: g.edges[e]["random_weight"] for e in g.edges}))
        return t

    def rec_choice(seq):
        x = random.choice(seq)
        log.append(("root", x))
        return x

    def rec_cut_choice(populated_graph, region_surcharge, cut_edge_list):
        c = _region_preferred_max_weight_choice(populated_graph, region_surcharge, cut_edge_list)
        log.append(("cut", c.edge, [x.edge for x in cut_edge_list]))
        return c

    method = partial(bipartition_tree, max_attempts=10000, spanning_tree_fn=rec_mst,
                     choice=rec_choice, cut_choice=rec_cut_choice)
    total = int(csr.pop.sum())
    random.seed(seed)
    plan = recursive_tree_part(graph, range(k), total / k, pop_col, eps, method=method)
    # split the log into one record per split: ... tree root [tree root ...] cut
    splits, cur = [], {"trees": [], "roots": []}
    for ev in log:
        if ev[0] == "tree":
            ids = np.fromiter((edge_id(e) for e in ev[1]), dtype=np.int64)
            vals = np.fromiter(ev[1].values(), dtype=np.float64)
            order = np.argsort(ids)
            cur["trees"].append((ids[order].astype(np.int32), vals[order]))
        elif ev[0] == "root":
            cur["roots"].append(index[ev[1]])
        else:
            cur["cut_edge"] = edge_id(ev[1])
            cur["n_cuts"] = len(ev[2])
            splits.append(cur)
            cur = {"trees": [], "roots": []}
    assert len(splits) == k - 1 and not cur["trees"]
    tree_ptr = np.zeros(len(splits) + 1, dtype=np.int64)
    w_ptr, ids, vals = [0], [], []
    for i, sp in enumerate(splits):
        tree_ptr[i + 1] = tree_ptr[i] + len(sp["trees"])
        for (i_, v_) in sp["trees"]:
            ids.append(i_)
            vals.append(v_)
            w_ptr.append(w_ptr[-1] + len(i_))
    np.savez_compressed(
        path,
        meta=json.dumps({"k": k, "epsilon": eps, "pop_target": total / k, "seed": seed,
                         "generator": "oracle/record_seed_plans.py"}),
        edge_uv=csr.edge_uv, pop=csr.pop,
        tree_ptr=tree_ptr, root=np.array([sp["roots"][-1] for sp in splits], dtype=np.int32),
        cut_edge=np.array([sp["cut_edge"] for sp in splits], dtype=np.int32),
        n_cuts=np.array([sp["n_cuts"] for sp in splits], dtype=np.int32),
        w_ptr=np.array(w_ptr, dtype=np.int64), w_eid=np.concatenate(ids), w_val=np.concatenate(vals),
        plan=np.array([plan[v] for v in csr.node_labels], dtype=np.uint8))
    print(f"wrote {path}: {k} districts, {int(tree_ptr[-1])} trees, {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    from gerrychain_b200 import synthetic as syn

    gold = os.path.join(REPO, "tests", "golden")
    record(syn.grid_graph(12, 12), 6, 0.1, "population", os.path.join(gold, "seedplan_grid12_k6.npz"), seed=3)
    record(syn.grid_graph(20, 20), 8, 0.05, "population", os.path.join(gold, "seedplan_grid20_k8.npz"), seed=4)
    record(syn.delaunay_graph(600, seed=0), 5, 0.05, "population",
           os.path.join(gold, "seedplan_delaunay600_k5.npz"), seed=5)
    record(syn.grid_graph(6, 6), 2, 0.2, "population", os.path.join(gold, "seedplan_grid6_k2.npz"), seed=6)


if __name__ == "__main__":
    main()
