#!/usr/bin/env python
"""Record the reference's own ``reversible_recom`` calls as replay fixtures
(tests/golden/reversible_*.npz).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/record_reversible.py

The UNMODIFIED reference is imported through the geopandas type-hint stub (oracle/refshim) and
its injection points are wrapped - no reference edits
(/root/reference/gerrychain/proposals/tree_proposals.py:149-267):

* the ``random`` module seen by tree_proposals.py -> shim recording ``random.choice(dist_pairs)``
  (:224), ``random.choice(list(pair_edges))`` (:230) and ``random.random()`` (:262);
* ``tree_proposals.uniform_spanning_tree`` (looked up when the call builds its partial, :211-216)
  -> wrapper recording the Wilson root and every successor draw per node (tree.py:112, :119);
* ``balance_edge_fn=`` -> find_balanced_edge_cuts_memoization with a recording ``choice`` (the
  BFS root, tree.py:377);
* ``choice=`` -> the chosen ``Cut.edge`` and the list of balanced cuts (:244).

Replay is by identity (district pair, edge id, node -> draw stack, root node, cut edge, the
acceptance draw), never by RNG stream position.  The oracle replays the records bit-exactly
(tests/test_oracle_vs_reference.py::test_reversible_replay_matches_reference); the device is
bit-exact with the oracle on identical Philox draws (tests/test_gpu_parity.py).
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
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

from record_reference import Recorder, import_reference  # noqa: E402


class _Shim:
    def __init__(self, log):
        self._log = log

    def choice(self, seq):
        x = random.choice(seq)
        self._log.append(("choice", x))
        return x

    def random(self):
        u = random.random()
        self._log.append(("u", u))
        return u

    def __getattr__(self, name):
        return getattr(random, name)


class ReversibleRecorder:
    def __init__(self, nx_graph, assignment, pop_col, k, epsilon, M):
        import_reference()
        from gerrychain import Graph, Partition
        from gerrychain.proposals import tree_proposals
        from gerrychain.tree import find_balanced_edge_cuts_memoization, uniform_spanning_tree
        from gerrychain.updaters import Tally, cut_edges
        from gerrychain_b200.csr import assignment_vector, csr_from_networkx

        self.graph = Graph.from_networkx(nx_graph)
        self.csr = csr_from_networkx(self.graph, pop_col)
        self.index = self.csr.node_index()
        self.eid = self.csr.edge_id_lookup()
        vec, parts = assignment_vector(self.csr, assignment)
        assert parts == list(range(k)), "fixtures use district labels 0..k-1"
        self.init_assignment = vec
        self.k, self.epsilon, self.M = k, float(epsilon), int(M)
        self.pop_target = int(self.csr.pop.sum()) / k
        self.log = []
        log = self.log

        def rec_wilson(graph):
            draws = []

            def ch(seq):
                x = random.choice(seq)
                draws.append((tuple(seq), x))
                return x

            t = uniform_spanning_tree(graph, choice=ch)
            log.append(("wilson", draws))
            return t

        def rec_root(seq):
            x = random.choice(seq)
            log.append(("root", x))
            return x

        def rec_balance(h, **kw):
            kw["choice"] = rec_root
            return find_balanced_edge_cuts_memoization(h, **kw)

        def rec_cut(cuts):
            c = random.choice(cuts)
            log.append(("cut", c.edge, [x.edge for x in cuts]))
            return c

        self._tp = tree_proposals
        self._wilson = rec_wilson
        self._shim = _Shim(log)
        self.proposal = partial(tree_proposals.reversible_recom, pop_col=pop_col, pop_target=self.pop_target,
                                epsilon=self.epsilon, balance_edge_fn=rec_balance, M=self.M, choice=rec_cut)
        self.state = Partition(self.graph, assignment,
                               {"population": Tally(pop_col, alias="population"), "cut_edges": cut_edges})

    def _edge_id(self, e):
        return self.eid[tuple(sorted((self.index[e[0]], self.index[e[1]])))]

    def step(self):
        del self.log[:]
        tp = self._tp
        saved = (tp.random, tp.uniform_spanning_tree)
        tp.random, tp.uniform_spanning_tree = self._shim, self._wilson
        try:
            new = self.proposal(self.state)
        finally:
            tp.random, tp.uniform_spanning_tree = saved
        rec = dict(d_out=-1, d_in=-1, pair_edge=-1, root=-1, cut_edge=-1, n_balanced=-1, accept_u=np.nan,
                   tree=None, outcome=0)
        choices = [ev[1] for ev in self.log if ev[0] == "choice"]
        rec["d_out"], rec["d_in"] = choices[0]
        if len(choices) > 1:
            e = choices[1]
            rec["pair_edge"] = self._edge_id(e)
            # parts_to_merge[0] is the district of edge[0] (:231-234); the oracle and the kernel take
            # the lower-indexed endpoint of the edge id: the same node as long as the reference lists
            # its edges (lower index, higher index), which holds for graphs built in node order
            assert self.index[e[0]] < self.index[e[1]], "edge orientation differs from (min, max)"
        for ev in self.log:
            if ev[0] == "wilson":
                rec["tree"] = Recorder._wilson_record(self, ev[1])
                rec["outcome"] = 1
                rec["n_balanced"] = 0
            elif ev[0] == "root":
                rec["root"] = self.index[ev[1]]
            elif ev[0] == "cut":
                rec["cut_edge"] = self._edge_id(ev[1])
                rec["n_balanced"] = len(ev[2])
            elif ev[0] == "u":
                rec["accept_u"] = ev[1]
                rec["outcome"] = 3 if new is not self.state else 2
        if new is not self.state:
            self.state.parent = None
            self.state = new
        m = self.state.assignment.mapping
        rec["assignment"] = np.fromiter((m[v] for v in self.csr.node_labels), dtype=np.uint8,
                                        count=self.csr.n_nodes)
        rec["n_cut_edges"] = len(self.state["cut_edges"])
        return rec

    def record(self, n_steps, path, seed):
        random.seed(seed)
        recs = [self.step() for _ in range(n_steps)]
        trees = [r["tree"] for r in recs if r["tree"] is not None]
        tree_idx, t = [], 0
        for r in recs:
            tree_idx.append(t if r["tree"] is not None else -1)
            t += r["tree"] is not None
        roots, nptr, nodes, sptr, svals = [], [0], [], [0], []
        for (_, wr, nd, ptr, vals) in trees:
            roots.append(wr)
            nodes.append(nd)
            base = sptr[-1]
            sptr.extend((ptr[1:] + base).tolist())
            svals.append(vals)
            nptr.append(nptr[-1] + len(nd))
        out = {
            "meta": json.dumps({"k": self.k, "epsilon": self.epsilon, "pop_target": self.pop_target, "M": self.M,
                                "seed": seed, "n_steps": n_steps, "generator": "oracle/record_reversible.py"}),
            "edge_uv": self.csr.edge_uv, "pop": self.csr.pop,
            "init_assignment": self.init_assignment.astype(np.uint8),
            "d_out": np.array([r["d_out"] for r in recs], dtype=np.int32),
            "d_in": np.array([r["d_in"] for r in recs], dtype=np.int32),
            "pair_edge": np.array([r["pair_edge"] for r in recs], dtype=np.int32),
            "root": np.array([r["root"] for r in recs], dtype=np.int32),
            "cut_edge": np.array([r["cut_edge"] for r in recs], dtype=np.int32),
            "n_balanced": np.array([r["n_balanced"] for r in recs], dtype=np.int32),
            "accept_u": np.array([r["accept_u"] for r in recs], dtype=np.float64),
            "outcome": np.array([r["outcome"] for r in recs], dtype=np.int32),
            "tree_idx": np.array(tree_idx, dtype=np.int64),
            "assignment": np.stack([r["assignment"] for r in recs]),
            "n_cut_edges": np.array([r["n_cut_edges"] for r in recs], dtype=np.int32),
            "wilson_root": np.array(roots, dtype=np.int32),
            "wn_ptr": np.array(nptr, dtype=np.int64),
            "wn_node": np.concatenate(nodes),
            "ws_ptr": np.array(sptr, dtype=np.int64),
            "ws_val": np.concatenate(svals),
        }
        np.savez_compressed(path, **out)
        oc = np.bincount(out["outcome"], minlength=4)
        print(f"wrote {path}: {n_steps} calls ({oc[0]} without a pair, {oc[1]} without a cut, {oc[2]} rejected, "
              f"{oc[3]} moves), {len(trees)} trees, {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    from gerrychain_b200 import synthetic as syn

    gold = os.path.join(REPO, "tests", "golden")
    from gerrychain.proposals.tree_proposals import ReversibilityError  # noqa: F401  (import_reference ran)

    g8 = syn.grid_graph(8, 8)
    ReversibleRecorder(g8, syn.grid_quadrants(8, 8), "population", 4, 0.15, M=8).record(
        900, os.path.join(gold, "reversible_grid8.npz"), seed=31)
    g12 = syn.grid_graph(12, 12)
    ReversibleRecorder(g12, syn.grid_quadrants(12, 12), "population", 4, 0.1, M=12).record(
        350, os.path.join(gold, "reversible_grid12.npz"), seed=32)


if __name__ == "__main__":
    import_reference()
    main()
