#!/usr/bin/env python
"""Record the reference's per-plan statistics as fixtures (tests/golden/stats_*.npz).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/record_statistics.py

For a list of plans on a graph with region columns it stores what the UNMODIFIED reference
reports for each plan:

* ``total_reg_splits(partition, col)``              updaters/county_splits.py:145-161
* ``compute_county_splits(partition, col, ...)``    updaters/county_splits.py:58-117
  (counties whose split is not NOT_SPLIT, and the sum of ``len(contains)``)
* ``contiguous(partition)``                         constraints/contiguity.py:169-181
* the number of parts of ``contiguous_components``  constraints/contiguity.py:226-241

The plans are states of a reference ReCom chain (contiguous by construction) and the same
states with random single-node reassignments (mostly not contiguous).
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

from record_reference import REFERENCE, import_reference  # noqa: E402


def build_cases():
    import networkx as nx

    from gerrychain_b200 import synthetic as syn

    cases = []
    # the reference's own region-aware test graph (tests/test_region_aware.py:37)
    with open(os.path.join(REFERENCE, "tests", "graphs_for_test", "8x8_with_muni.json")) as f:
        data = json.load(f)
    g = nx.readwrite.json_graph.adjacency_graph(data)
    cases.append(dict(name="8x8_muni", graph=g, pop_col="TOTPOP", seed_col="district",
                      region_cols=["muni", "county", "county2"], eps=0.1, steps=30, k=None))
    # a Delaunay graph with Voronoi counties (BASELINE.json config5 in small)
    g2 = syn.voronoi_counties(syn.delaunay_graph(600, seed=3, pop_range=(50, 250)), 12, seed=5)
    cases.append(dict(name="delaunay600", graph=g2, pop_col="population", seed_col=None,
                      region_cols=["county"], eps=0.1, steps=30, k=6))
    return cases


def record(case):
    gc = import_reference()
    from gerrychain import Graph, MarkovChain, Partition
    from gerrychain.accept import always_accept
    from gerrychain.constraints import contiguous, within_percent_of_ideal_population
    from gerrychain.constraints.contiguity import contiguous_components
    from gerrychain.proposals import recom
    from gerrychain.tree import recursive_tree_part
    from gerrychain.updaters import Tally, cut_edges
    from gerrychain.updaters.county_splits import CountySplit, compute_county_splits, total_reg_splits

    random.seed(20240 + len(case["name"]))
    graph = Graph.from_networkx(case["graph"])
    labels = list(graph.nodes)
    index = {lab: i for i, lab in enumerate(labels)}
    pop_col = case["pop_col"]
    updaters = {"population": Tally(pop_col, alias="population"), "cut_edges": cut_edges}
    if case["seed_col"]:
        initial = Partition(graph, case["seed_col"], updaters)
    else:
        total = sum(graph.nodes[v][pop_col] for v in graph.nodes)
        assignment = recursive_tree_part(graph, range(case["k"]), total / case["k"], pop_col, case["eps"])
        initial = Partition(graph, assignment, updaters)
    parts = sorted(initial.parts)
    rank = {p: i for i, p in enumerate(parts)}
    k = len(parts)
    ideal = sum(initial["population"].values()) / k
    chain = MarkovChain(partial(recom, pop_col=pop_col, pop_target=ideal, epsilon=case["eps"], node_repeats=1),
                        [within_percent_of_ideal_population(initial, case["eps"])], always_accept, initial,
                        total_steps=case["steps"])
    plans = []
    for state in chain:
        vec = [rank[state.assignment.mapping[lab]] for lab in labels]
        plans.append(vec)
        for n_flip in (1, 3, 8):  # perturbed copies: random nodes to random districts
            v2 = list(vec)
            for _ in range(n_flip):
                v2[random.randrange(len(v2))] = random.randrange(k)
            plans.append(v2)
    plans = np.array(plans, dtype=np.uint8)

    region_ids, n_regions = [], []
    for col in case["region_cols"]:
        codes = {}
        region_ids.append([codes.setdefault(graph.nodes[lab][col], len(codes)) for lab in labels])
        n_regions.append(len(codes))
    region_ids = np.array(region_ids, dtype=np.int32)

    # an Election over two synthetic vote columns (updaters/election.py:7-166)
    from gerrychain.updaters import Election

    vrng = np.random.default_rng(11)
    votes = vrng.integers(0, 400, (2, len(labels))).astype(np.int32)
    for j, lab in enumerate(labels):
        graph.nodes[lab]["votes_a"] = int(votes[0, j])
        graph.nodes[lab]["votes_b"] = int(votes[1, j])
    election = Election("synthetic", {"A": "votes_a", "B": "votes_b"}, alias="election")
    updaters = dict(updaters, election=election)
    vote_counts = np.zeros((len(plans), 2, k), dtype=np.int64)
    seats = np.zeros((len(plans), 2), dtype=np.int32)

    edge_splits = np.zeros((len(plans), len(case["region_cols"])), dtype=np.int32)
    split_regions = np.zeros_like(edge_splits)
    pieces = np.zeros_like(edge_splits)
    is_contiguous = np.zeros(len(plans), dtype=np.uint8)
    district_pieces = np.zeros((len(plans), k), dtype=np.int32)
    for i, vec in enumerate(plans):
        part = Partition(graph, {lab: int(vec[index[lab]]) for lab in labels}, updaters)
        for c, col in enumerate(case["region_cols"]):
            edge_splits[i, c] = total_reg_splits(part, col)
            info = compute_county_splits(part, col, "unused")
            split_regions[i, c] = sum(1 for ci in info.values() if ci.split != CountySplit.NOT_SPLIT)
            pieces[i, c] = sum(len(ci.contains) for ci in info.values())
        results = part["election"]
        for pi, party in enumerate(("A", "B")):
            vote_counts[i, pi] = [results.count(party, p_) if p_ in part.parts else 0 for p_ in range(k)]
            seats[i, pi] = results.seats(party)
        is_contiguous[i] = bool(contiguous(part))
        comps = contiguous_components(part)
        for part_id, subgraphs in comps.items():
            district_pieces[i, part_id] = len(subgraphs)
    edges = np.array([(index[u], index[v]) for u, v in graph.edges], dtype=np.int32)
    pop = np.array([graph.nodes[lab][pop_col] for lab in labels], dtype=np.int32)
    meta = dict(name=case["name"], k=k, region_cols=case["region_cols"], n_regions=n_regions,
                reference="mggg/GerryChain at /root/reference, run unmodified")
    out = os.path.join(REPO, "tests", "golden", f"stats_{case['name']}.npz")
    np.savez_compressed(out, meta=json.dumps(meta), edge_uv=edges, pop=pop, region_ids=region_ids, plans=plans,
                        edge_splits=edge_splits, split_regions=split_regions, pieces=pieces,
                        is_contiguous=is_contiguous, district_pieces=district_pieces, votes=votes,
                        vote_counts=vote_counts, seats=seats)
    print(out, "plans", len(plans), "contiguous", int(is_contiguous.sum()), "edge_splits mean", edge_splits.mean(0))


if __name__ == "__main__":
    for case in build_cases():
        record(case)
