#!/usr/bin/env python
"""Record the reference's own ReCom steps as replay fixtures (tests/golden/*.npz).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/record_reference.py            # regenerates every fixture

It imports the UNMODIFIED reference through a geopandas type-hint stub
(oracle/refshim) and wraps its public injection points - no reference edits:

* ``gerrychain.proposals.tree_proposals.random`` -> shim whose ``choice`` records the
  cut edge picked at /root/reference/gerrychain/proposals/tree_proposals.py:106;
* ``spanning_tree_fn`` wrapper: after ``random_spanning_tree`` it reads the
  ``random_weight`` the reference leaves on every subgraph edge
  (/root/reference/gerrychain/tree.py:89);  for Wilson, a ``choice`` wrapper records
  the root (tree.py:112) and every successor draw per node (tree.py:119);
* ``choice`` wrapper: the BFS root (tree.py:377);
* ``cut_choice`` wrapper with the parameter names that keep ``is_region_cut`` true
  (tree.py:685-688): the balanced cut list and the chosen ``Cut.edge``.

Replay is by identity (edge id -> weight, node -> draw stack, root node, cut edge),
never by RNG stream position (SURVEY.md §8c determinism caveats).
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
REFERENCE = os.environ.get("GERRYCHAIN_REFERENCE", "/root/reference")


def import_reference():
    """Import the reference package from /root/reference behind the geopandas stub."""
    for p in (REFERENCE, os.path.join(HERE, "refshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gerrychain  # noqa: F401

    return gerrychain


class _RandomShim:
    """Stands in for the ``random`` module inside tree_proposals.py only."""

    def __init__(self, log):
        self._log = log

    def choice(self, seq):
        picked = random.choice(seq)
        self._log.append(("pair_edge", picked))
        return picked

    def random(self):
        return random.random()

    def __getattr__(self, name):
        return getattr(random, name)


class Recorder:
    """Runs reference ``recom`` steps and emits one replay record per step."""

    def __init__(self, nx_graph, assignment, pop_col, k, epsilon, node_repeats=1,
                 region_surcharge=None, tree="mst"):
        gerrychain = import_reference()
        from gerrychain import Graph, Partition
        from gerrychain.proposals import tree_proposals
        from gerrychain.tree import (
            bipartition_tree,
            random_spanning_tree,
            uniform_spanning_tree,
            _region_preferred_max_weight_choice,
        )
        from gerrychain.updaters import Tally, cut_edges

        sys.path.insert(0, REPO)
        from gerrychain_b200.csr import assignment_vector, csr_from_networkx

        self.tree_kind = tree
        self.graph = Graph.from_networkx(nx_graph)
        self.csr = csr_from_networkx(self.graph, pop_col, region_surcharge)
        self.index = self.csr.node_index()
        self.eid = self.csr.edge_id_lookup()
        vec, parts = assignment_vector(self.csr, assignment)
        assert parts == list(range(k)), "fixtures use district labels 0..k-1"
        self.init_assignment = vec
        self.k = k
        self.epsilon = float(epsilon)
        self.node_repeats = int(node_repeats)
        self.pop_col = pop_col
        self.region_surcharge = region_surcharge
        total = int(self.csr.pop.sum())
        self.pop_target = total / k
        self.log = []
        log = self.log

        def rec_mst(graph, region_surcharge=None):
            t = random_spanning_tree(graph, region_surcharge=region_surcharge)
            w = {e: graph.edges[e]["random_weight"] for e in graph.edges}
            log.append(("tree_weights", w))
            return t

        def rec_wilson(graph):
            draws = []

            def ch(seq):
                x = random.choice(seq)
                draws.append((tuple(seq), x))
                return x

            t = uniform_spanning_tree(graph, choice=ch)
            # first draw is the root out of list(node_indices); the rest alternate
            # (current node is implied: it is the node whose neighbours were listed)
            log.append(("wilson", draws, t))
            return t

        def rec_choice(seq):
            x = random.choice(seq)
            log.append(("root", x))
            return x

        def rec_cut_choice(populated_graph, region_surcharge, cut_edge_list):
            c = _region_preferred_max_weight_choice(
                populated_graph, region_surcharge, cut_edge_list
            )
            log.append(("cut", c.edge, [x.edge for x in cut_edge_list]))
            return c

        method = partial(
            bipartition_tree,
            spanning_tree_fn=rec_mst if tree == "mst" else rec_wilson,
            choice=rec_choice,
            cut_choice=rec_cut_choice,
        )
        self.proposal = partial(
            tree_proposals.recom,
            pop_col=pop_col,
            pop_target=self.pop_target,
            epsilon=self.epsilon,
            node_repeats=self.node_repeats,
            region_surcharge=region_surcharge,
            method=method,
        )
        self._tp = tree_proposals
        self._shim = _RandomShim(log)
        self.state = Partition(
            self.graph,
            assignment,
            {"population": Tally(pop_col, alias="population"), "cut_edges": cut_edges},
        )

    # -- helpers -----------------------------------------------------------
    def _edge_id(self, e):
        return self.eid[tuple(sorted((self.index[e[0]], self.index[e[1]])))]

    def _assignment_vec(self, partition):
        m = partition.assignment.mapping
        labels = self.csr.node_labels
        return np.fromiter((m[v] for v in labels), dtype=np.int32, count=len(labels))

    def step(self):
        """One reference recom step -> dict record."""
        del self.log[:]
        saved = self._tp.random
        self._tp.random = self._shim
        try:
            new = self.proposal(self.state)
        finally:
            self._tp.random = saved
        self.state.parent = None  # chain.py:188-189
        self.state = new
        rec = {"trees": [], "roots": []}
        for ev in self.log:
            if ev[0] == "pair_edge":
                rec["pair_edge"] = self._edge_id(ev[1])
            elif ev[0] == "tree_weights":
                ids = np.fromiter((self._edge_id(e) for e in ev[1]), dtype=np.int64)
                vals = np.fromiter(ev[1].values(), dtype=np.float64)
                order = np.argsort(ids)
                rec["trees"].append((ids[order].astype(np.int32), vals[order]))
            elif ev[0] == "wilson":
                rec["trees"].append(self._wilson_record(ev[1]))
            elif ev[0] == "root":
                rec["roots"].append(self.index[ev[1]])
            elif ev[0] == "cut":
                rec["cut_edge"] = self._edge_id(ev[1])
                rec["balanced"] = sorted(self._edge_id(e) for e in ev[2])
        rec["assignment"] = self._assignment_vec(new)
        rec["n_cut_edges"] = len(new["cut_edges"])
        rec["population"] = np.array(
            [new["population"][d] for d in range(self.k)], dtype=np.int64
        )
        rec["cut_edge_ids"] = np.array(
            sorted(self._edge_id(e) for e in new["cut_edges"]), dtype=np.int32
        )
        return rec

    def _wilson_record(self, draws):
        """Per-node stacks of chosen neighbour node indices + the Wilson root."""
        (seq0, root) = draws[0]
        stacks = {}
        # Reconstruct which node each draw belongs to by re-walking the reference's
        # loop (tree.py:116-125) with the recorded choices.
        it = iter(draws[1:])
        tree_nodes = {root}
        nxt = {root: None}
        order = list(seq0)  # == list(graph.node_indices) in the same process
        for node in order:
            u = node
            while u not in tree_nodes:
                seq, x = next(it)
                stacks.setdefault(u, []).append(x)
                nxt[u] = x
                u = x
            u = node
            while u not in tree_nodes:
                tree_nodes.add(u)
                u = nxt[u]
        assert next(it, None) is None
        nodes = np.array(sorted(self.index[v] for v in stacks), dtype=np.int32)
        inv = {self.index[v]: v for v in stacks}
        ptr = np.zeros(len(nodes) + 1, dtype=np.int32)
        vals = []
        for i, vi in enumerate(nodes):
            s = stacks[inv[int(vi)]]
            vals.extend(self.index[x] for x in s)
            ptr[i + 1] = len(vals)
        return ("wilson", self.index[root], nodes, ptr, np.array(vals, dtype=np.int32))

    # -- fixture writer ----------------------------------------------------
    def record(self, n_steps, path, seed):
        random.seed(seed)
        recs = [self.step() for _ in range(n_steps)]
        out = {
            "meta": json.dumps(
                {
                    "k": self.k,
                    "epsilon": self.epsilon,
                    "pop_target": self.pop_target,
                    "node_repeats": self.node_repeats,
                    "tree": self.tree_kind,
                    "seed": seed,
                    "n_steps": n_steps,
                    "region_names": list(self.csr.region_names),
                    "generator": "oracle/record_reference.py",
                }
            ),
            "edge_uv": self.csr.edge_uv,
            "pop": self.csr.pop,
            "region_ids": self.csr.region_ids,
            "surcharges": self.csr.surcharges,
            "init_assignment": self.init_assignment.astype(np.uint8),
            "pair_edge": np.array([r["pair_edge"] for r in recs], dtype=np.int32),
            "root": np.array([r["roots"][-1] for r in recs], dtype=np.int32),
            "n_roots": np.array([len(r["roots"]) for r in recs], dtype=np.int32),
            "cut_edge": np.array([r["cut_edge"] for r in recs], dtype=np.int32),
            "assignment": np.stack([r["assignment"] for r in recs]).astype(np.uint8),
            "n_cut_edges": np.array([r["n_cut_edges"] for r in recs], dtype=np.int32),
            "population": np.stack([r["population"] for r in recs]),
        }
        bal_ptr = np.zeros(n_steps + 1, dtype=np.int32)
        bal = []
        for i, r in enumerate(recs):
            bal.extend(r["balanced"])
            bal_ptr[i + 1] = len(bal)
        out["balanced_ptr"] = bal_ptr
        out["balanced"] = np.array(bal, dtype=np.int32)
        tree_ptr = np.zeros(n_steps + 1, dtype=np.int32)
        n_trees = 0
        for i, r in enumerate(recs):
            n_trees += len(r["trees"])
            tree_ptr[i + 1] = n_trees
        out["tree_ptr"] = tree_ptr
        if self.tree_kind == "mst":
            w_ptr = np.zeros(n_trees + 1, dtype=np.int64)
            ids, vals = [], []
            t = 0
            for r in recs:
                for (i_, v_) in r["trees"]:
                    ids.append(i_)
                    vals.append(v_)
                    w_ptr[t + 1] = w_ptr[t] + len(i_)
                    t += 1
            out["w_ptr"] = w_ptr
            out["w_eid"] = np.concatenate(ids)
            out["w_val"] = np.concatenate(vals)
        else:
            roots, nptr, nodes, sptr, svals = [], [0], [], [0], []
            for r in recs:
                for (_, wr, nd, ptr, vals) in r["trees"]:
                    roots.append(wr)
                    nodes.append(nd)
                    base = sptr[-1]
                    sptr.extend((ptr[1:] + base).tolist())
                    svals.append(vals)
                    nptr.append(nptr[-1] + len(nd))
            out["wilson_root"] = np.array(roots, dtype=np.int32)
            out["wn_ptr"] = np.array(nptr, dtype=np.int64)  # per tree -> node slice
            out["wn_node"] = np.concatenate(nodes)
            out["ws_ptr"] = np.array(sptr, dtype=np.int64)  # per (tree,node) -> draws
            out["ws_val"] = np.concatenate(svals)
        np.savez_compressed(path, **out)
        print(f"wrote {path}: {n_steps} steps, {n_trees} trees, "
              f"{os.path.getsize(path) / 1024:.0f} KiB")
        return recs


def main():
    sys.path.insert(0, REPO)
    from gerrychain_b200 import synthetic as syn

    gold = os.path.join(REPO, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    # C1 shape: Grid 20x20, quadrants, eps 0.05 (BASELINE.json configs[0])
    g = syn.grid_graph(20, 20)
    Recorder(g, syn.grid_quadrants(20, 20), "population", 4, 0.05).record(
        120, os.path.join(gold, "replay_grid20_mst.npz"), seed=2024)
    Recorder(g, syn.grid_quadrants(20, 20), "population", 4, 0.05, tree="wilson").record(
        60, os.path.join(gold, "replay_grid20_wilson.npz"), seed=7)
    # node_repeats = 3 exercises the "same tree, new root" retries (tree.py:680-682)
    Recorder(g, syn.grid_quadrants(20, 20), "population", 4, 0.05, node_repeats=3).record(
        40, os.path.join(gold, "replay_grid20_mst_nr3.npz"), seed=5)

    # with_diagonals grid of tests/updaters/test_cut_edges.py:128-161
    gd = syn.grid_graph(10, 10, with_diagonals=True)
    Recorder(gd, syn.grid_quadrants(10, 10), "population", 4, 0.1).record(
        100, os.path.join(gold, "replay_grid10diag_mst.npz"), seed=11)

    # C2 shape (short): Grid 100x100, 10 stripes
    g2 = syn.grid_graph(100, 100)
    Recorder(g2, syn.grid_stripes(100, 100, 10), "population", 10, 0.05).record(
        12, os.path.join(gold, "replay_grid100_mst.npz"), seed=3)

    # C3 shape (short): Delaunay 2.4k nodes / 6 districts keeps ~800-node merges
    # while the fixture stays small; seed plan from the reference's recursive_tree_part
    import_reference()
    from gerrychain import Graph
    from gerrychain.tree import recursive_tree_part

    g3 = syn.delaunay_graph(2400, seed=0)
    random.seed(0)
    tot = sum(g3.nodes[v]["population"] for v in g3)
    plan = recursive_tree_part(Graph.from_networkx(g3), range(6), tot / 6, "population", 0.02)
    Recorder(g3, plan, "population", 6, 0.02).record(
        25, os.path.join(gold, "replay_delaunay2400_mst.npz"), seed=1)
    Recorder(g3, plan, "population", 6, 0.02, tree="wilson").record(
        10, os.path.join(gold, "replay_delaunay2400_wilson.npz"), seed=2)

    # C5 shape: region-aware, counties + surcharge
    syn.voronoi_counties(g3, 16, seed=1, key="county")
    Recorder(g3, plan, "population", 6, 0.02, region_surcharge={"county": 0.5}).record(
        25, os.path.join(gold, "replay_delaunay2400_region.npz"), seed=4)

    # reference fixture 8x8_with_muni.json: two region columns (power-set choice)
    with open(os.path.join(REFERENCE, "tests/graphs_for_test/8x8_with_muni.json")) as f:
        data = json.load(f)
    import networkx
    g8 = networkx.readwrite.json_graph.adjacency_graph(data)
    plan8 = {v: int(g8.nodes[v]["district"]) - 1 for v in g8}
    Recorder(g8, plan8, "TOTPOP", 8, 0.25,
             region_surcharge={"muni": 1.0, "county": 0.4}).record(
        80, os.path.join(gold, "replay_8x8_region2.npz"), seed=9)
    Recorder(g8, plan8, "TOTPOP", 8, 0.25,
             region_surcharge={"muni": 0.3, "county": 0.2, "county2": 0.7}).record(
        80, os.path.join(gold, "replay_8x8_region3.npz"), seed=10)


if __name__ == "__main__":
    main()
