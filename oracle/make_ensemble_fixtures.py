#!/usr/bin/env python
"""Ensemble fixtures for the KS tests (tests/golden/ensemble_*.npz).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/make_ensemble_fixtures.py

Runs the UNMODIFIED reference (`MarkovChain` + `recom`, one chain per process-pool worker,
`random.seed(chain id)` - the reference's own pattern for ensembles,
/root/reference/tests/test_region_aware.py:82-99) for many independent chains from one seed
plan and records, at a fixed step index, the statistics the north star names: the number of
cut edges and the population deviation of every district (constraints/validity.py:96-122).
Samples are taken ACROSS chains at a fixed step (independent), never along a chain.
"""

from __future__ import annotations

import json
import os
import random
import sys
from concurrent.futures import ProcessPoolExecutor
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle.record_reference import import_reference  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")

CASES = {
    # name: (grid dims, k via plan, epsilon, tree, region_surcharge, steps, chains)
    "grid20_mst": dict(dims=(20, 20), eps=0.05, tree="mst", steps=60, chains=384),
    "grid20_wilson": dict(dims=(20, 20), eps=0.05, tree="wilson", steps=60, chains=384),
    "grid12_region": dict(dims=(12, 12), eps=0.1, tree="mst", steps=60, chains=384, region=True),
    # Metropolis acceptance on the number of cut edges (accept.py:19-35)
    "grid20_cutaccept": dict(dims=(20, 20), eps=0.05, tree="mst", steps=60, chains=384, accept="cut_edge"),
    # reversible ReCom (tree_proposals.py:149-267): most proposals are self-loops, hence many steps
    # random planar graph with real-valued (integer) populations: 600-node Delaunay, 5 districts
    "delaunay600_mst": dict(graph="delaunay", n=600, k=5, eps=0.05, tree="mst", steps=60, chains=384),
    "grid12_reversible": dict(dims=(12, 12), eps=0.1, tree="wilson", steps=1500, chains=256, reversible_M=12),
    # BASELINE configs 3 and 5 at their stated size: Delaunay 9k nodes, 18 districts, eps 0.02, the
    # committed seed plan (tests/golden/plan_delaunay9000_k18.npy); config 5 adds 67 Voronoi
    # counties with region_surcharge {"county": 0.5} (the reference's own region-aware pattern,
    # /root/reference/tests/test_region_aware.py:82-99)
    "delaunay9000_region": dict(graph="delaunay", n=9000, k=18, eps=0.02, tree="mst", steps=50, chains=256,
                                region=True, counties=67, plan_file="plan_delaunay9000_k18.npy"),
    "delaunay9000_mst": dict(graph="delaunay", n=9000, k=18, eps=0.02, tree="mst", steps=50, chains=256,
                             plan_file="plan_delaunay9000_k18.npy"),
    "delaunay9000_wilson": dict(graph="delaunay", n=9000, k=18, eps=0.02, tree="wilson", steps=50, chains=192,
                                plan_file="plan_delaunay9000_k18.npy"),
    # BASELINE config 4 at its stated size: Delaunay 200k nodes, 50 districts, eps 0.05,
    # uniform_spanning_tree, the committed seed plan.  ~1.7 s per reference step: each worker
    # process builds the graph once and runs a block of chains (chains_per_task)
    "delaunay200000_wilson": dict(graph="delaunay", n=200000, k=50, eps=0.05, tree="wilson", steps=20, chains=128,
                                  plan_file="plan_delaunay200000_k50.npy", pop_range=[20, 200], chains_per_task=16),
}


def _build(case):
    import_reference()
    import networkx
    from gerrychain import Graph, Partition
    from gerrychain.updaters import Tally, cut_edges

    if case.get("graph") == "delaunay":
        from gerrychain.tree import recursive_tree_part
        from gerrychain_b200 import synthetic as syn

        g = syn.delaunay_graph(case["n"], seed=0, **({"pop_range": tuple(case["pop_range"])} if case.get("pop_range") else {}))
        if case.get("counties"):
            syn.voronoi_counties(g, case["counties"], seed=1, key="county")
        graph = Graph.from_networkx(g)
        if case.get("plan_file"):
            vec = np.load(os.path.join(GOLDEN, case["plan_file"]))
            plan = {v: int(d) for v, d in zip(g.nodes, vec)}
        else:
            random.seed(12345)  # one fixed seed plan for every chain of the ensemble
            tot = sum(g.nodes[v]["population"] for v in g)
            plan = recursive_tree_part(graph, range(case["k"]), tot / case["k"], "population", case["eps"])
        part = Partition(graph, plan, {"population": Tally("population"), "cut_edges": cut_edges})
        return graph, part, plan
    m, n = case["dims"]
    g = networkx.grid_2d_graph(m, n)
    for (i, j) in g.nodes:
        g.nodes[(i, j)]["population"] = 1
        g.nodes[(i, j)]["county"] = (i // (m // 3)) * 3 + j // (n // 3)
    plan = {(i, j): (0 if i < m // 2 else 1) + (0 if j < n // 2 else 2) for (i, j) in g.nodes}
    graph = Graph.from_networkx(g)
    part = Partition(graph, plan, {"population": Tally("population"), "cut_edges": cut_edges})
    return graph, part, plan


def run_chain_block(first_seed, case):
    """chains first_seed .. first_seed + chains_per_task - 1 on ONE build of the graph (big graphs)."""
    built = _build(case)
    return [run_chain(seed, case, built) for seed in range(first_seed, first_seed + case["chains_per_task"])]


def run_chain(seed, case, built=None):
    import_reference()
    from gerrychain import MarkovChain, accept, constraints
    from gerrychain.proposals import recom
    from gerrychain.tree import bipartition_tree, uniform_spanning_tree

    graph, part, plan = built if built is not None else _build(case)
    random.seed(seed)
    ideal = sum(part["population"].values()) / len(part)
    method = bipartition_tree
    if case["tree"] == "wilson":
        method = partial(bipartition_tree, spanning_tree_fn=uniform_spanning_tree)
    kw = dict(pop_col="population", pop_target=ideal, epsilon=case["eps"], node_repeats=1, method=method)
    if case.get("region"):
        kw["region_surcharge"] = {"county": 0.5}
    proposal = partial(recom, **kw)
    if case.get("reversible_M"):
        from gerrychain.proposals.tree_proposals import reversible_recom

        proposal = partial(reversible_recom, pop_col="population", pop_target=ideal, epsilon=case["eps"],
                           M=case["reversible_M"])
    chain = MarkovChain(proposal,
                        [constraints.within_percent_of_ideal_population(part, case["eps"])],
                        accept.cut_edge_accept if case.get("accept") == "cut_edge" else accept.always_accept,
                        part, total_steps=case["steps"] + 1)
    state = None
    for state in chain:
        pass
    pops = [state["population"][d] for d in sorted(state["population"])]
    county_splits = 0
    if case.get("region"):
        for nodes in state.parts.values():
            county_splits += len({graph.nodes[v]["county"] for v in nodes}) - 1
    return len(state["cut_edges"]), pops, county_splits


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
            if case.get("chains_per_task"):
                blocks = ex.map(partial(run_chain_block, case=case), range(0, case["chains"], case["chains_per_task"]))
                rows = [r for b in blocks for r in b]
            else:
                rows = list(ex.map(partial(run_chain, case=case), range(case["chains"])))
        _, part, plan = _build(case)
        labels = list(part.graph.nodes)
        out = os.path.join(GOLDEN, f"ensemble_{name}.npz")
        np.savez_compressed(
            out, meta=json.dumps(case), cut_edges=np.array([r[0] for r in rows], dtype=np.int32),
            population=np.array([r[1] for r in rows], dtype=np.int64),
            county_splits=np.array([r[2] for r in rows], dtype=np.int32),
            plan=np.array([plan[v] for v in labels], dtype=np.uint8))
        print(out, "mean cut edges", np.mean([r[0] for r in rows]))


if __name__ == "__main__":
    main()
