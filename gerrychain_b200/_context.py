"""Per-graph host cache behind the GerryChain-shaped API: node index, canonical CSR per
population column, and the device engines built for it.

The reference keeps all of this implicitly in networkx dicts and ``functools.lru_cache``s
(/root/reference/gerrychain/graph/graph.py:508-577); here it is the bridge from a
``networkx``-style graph object to the flat arrays the C ABI takes.
"""

from __future__ import annotations

import random
import weakref
from typing import Dict, Hashable, Tuple

import numpy as np

from .config import EngineConfig
from .csr import GraphCSR, csr_from_edges

_contexts: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def unwrap_graph(graph):
    """Accept a networkx graph or anything forwarding ``.nodes`` / ``.edges`` to one
    (the reference's FrozenGraph, graph/graph.py:508)."""
    inner = getattr(graph, "graph", None)
    if inner is not None and hasattr(inner, "nodes") and not isinstance(inner, dict):
        return inner
    return graph


def context_of(graph) -> "GraphContext":
    g = unwrap_graph(graph)
    ctx = _contexts.get(g)
    if ctx is None:
        ctx = GraphContext(g)
        _contexts[g] = ctx
    return ctx


class GraphContext:
    def __init__(self, graph):
        if not hasattr(graph, "nodes") or not hasattr(graph, "edges"):
            raise TypeError(f"Unsupported Graph object with type {type(graph)}")
        self.graph_ref = weakref.ref(graph)
        self.labels = list(graph.nodes)
        self.index: Dict[Hashable, int] = {lab: i for i, lab in enumerate(self.labels)}
        self.n_nodes = len(self.labels)
        self.edges = np.array([(self.index[u], self.index[v]) for u, v in graph.edges],
                              dtype=np.int64).reshape(-1, 2)
        self._csr: Dict[Tuple, GraphCSR] = {}
        self._engines: Dict[Tuple, object] = {}
        self._edge_labels = None
        self._region_codes: Dict[str, Tuple[np.ndarray, int]] = {}
        self._geometry = None
        self._region_members: Dict[str, Tuple] = {}

    @property
    def graph(self):
        g = self.graph_ref()
        if g is None:
            raise ReferenceError("the graph of this partition has been garbage collected")
        return g

    # -- flat arrays -----------------------------------------------------------------
    def column(self, fields) -> np.ndarray:
        """Integer node-attribute column (sum of ``fields``) in node-index order."""
        g = self.graph
        if fields is None:  # no attribute needed (cut edges only): unit weights
            return np.ones(self.n_nodes, dtype=np.int64)
        if isinstance(fields, str):
            fields = (fields,)
        col = np.zeros(self.n_nodes, dtype=np.float64)
        for f in fields:
            col += np.array([g.nodes[lab][f] for lab in self.labels], dtype=np.float64)
        if not np.all(col == np.floor(col)):
            raise TypeError(f"node attribute {fields!r} is not integer valued; the device "
                            "engine tallies integers only")
        return col.astype(np.int64)

    def float_column(self, fields):
        """The float64 column of ``fields`` when the device cannot tally it exactly (non-integer
        values, NaN, or |sum| >= 2^31), else None (integer column: device path)."""
        key = ("float", tuple(fields))
        if key not in self._region_codes:
            g = self.graph
            col = np.zeros(self.n_nodes, dtype=np.float64)
            for f in fields:
                col += np.array([g.nodes[lab][f] for lab in self.labels], dtype=np.float64)
            finite = col[~np.isnan(col)]
            exact = (not np.isnan(col).any()) and np.all(finite == np.floor(finite)) and np.abs(finite).sum() < 2**31
            self._region_codes[key] = (None if exact else col, 0)
        return self._region_codes[key][0]

    def csr(self, fields, region_surcharge=None) -> GraphCSR:
        if isinstance(fields, str):
            fields = (fields,)
        if fields is None and self._csr:  # any cached column serves when only the edges matter
            return next(iter(self._csr.values()))
        fields = tuple(fields) if fields is not None else None
        rkey = tuple(region_surcharge.items()) if region_surcharge else ()
        key = (fields, rkey)
        csr = self._csr.get(key)
        if csr is None:
            base = self._csr.get((fields, ()))
            if base is None:
                col = self.column(fields)
                if np.abs(col).sum() >= 2**31:
                    raise OverflowError(f"total of {fields!r} must fit in int32")
                base = csr_from_edges(self.n_nodes, self.edges, col, node_labels=self.labels)
                self._csr[(fields, ())] = base
            csr = base
            if region_surcharge:
                g = self.graph
                cols = []
                for name in region_surcharge:
                    codes: Dict = {}
                    col = np.empty(self.n_nodes, dtype=np.int32)
                    for i, lab in enumerate(self.labels):
                        val = g.nodes[lab].get(name)
                        col[i] = -1 if val is None else codes.setdefault(val, len(codes))
                    cols.append(col)
                csr = base.with_regions(np.stack(cols), list(region_surcharge.values()),
                                        tuple(region_surcharge))
                self._csr[key] = csr
        return csr

    def edge_labels(self, csr: GraphCSR):
        """Edge id -> the tuple the reference stores in ``cut_edges``
        (updaters/cut_edges.py:110-114: ``tuple(sorted(edge))`` over node labels)."""
        if self._edge_labels is None:
            out = []
            for u, v in csr.edge_uv:
                a, b = self.labels[u], self.labels[v]
                try:
                    out.append((a, b) if a <= b else (b, a))
                except TypeError:
                    out.append((a, b))
            self._edge_labels = out
        return self._edge_labels

    # -- engines ---------------------------------------------------------------------
    def engine(self, key: Tuple, fields, k: int, region_surcharge=None, **cfg_kw):
        """One single-chain device engine per distinct (column, configuration)."""
        if isinstance(fields, str):
            fields = (fields,)
        full = (key, tuple(fields) if fields is not None else None, k)
        eng = self._engines.get(full)
        if eng is None:
            from .engine import Engine

            csr = self.csr(fields, region_surcharge)
            cfg_kw.setdefault("pop_target", float(csr.pop.sum()) / k)
            cfg_kw.setdefault("epsilon", 1.0)
            # the reference seeds everything from Python's global `random`; so does this
            cfg_kw.setdefault("seed", random.getrandbits(64))
            cfg = EngineConfig(n_districts=k, n_chains=1, **cfg_kw)
            eng = Engine(csr, cfg, np.zeros(self.n_nodes, dtype=np.uint8))
            self._engines[full] = eng
        return eng

    def grow_engine(self, eng):
        """Replace a single-chain engine whose merged-pair caps were too small (RC_ERR_CAPACITY)
        by one with doubled caps - up to the whole graph, where the engine spills to its L2
        scratch instead of failing.  Same configuration and Philox key: the step that failed is
        simply taken again.  Returns the new engine, or None when the caps cannot grow."""
        from dataclasses import replace

        from .engine import Engine

        info = eng.info()
        csr, cfg = eng.csr, eng.cfg
        limit_n, limit_e = min(csr.n_nodes, 32760), min(csr.n_edges, 65535)
        if info.cap_nodes >= limit_n and info.cap_edges >= limit_e:
            return None
        new_cfg = replace(cfg, cap_nodes=min(2 * info.cap_nodes, limit_n), cap_edges=min(2 * info.cap_edges, limit_e))
        new = Engine(csr, new_cfg, np.zeros(self.n_nodes, dtype=np.uint8))
        new.proposal_no.copy_(eng.proposal_no)  # the Philox position of the chain carries over
        for key, val in list(self._engines.items()):
            if val is eng:
                self._engines[key] = new
        eng.close()
        return new

    def scan(self, vec: np.ndarray, fields, k: int):
        """First-time Tally + cut_edges of a plan (rc_engine_init_state):
        returns (tally[k] int64, cut_mask words uint32, cut_count)."""
        eng = self.engine(("scan",), fields, k)
        eng.set_assignment(vec)
        eng.synchronize()
        return (eng.tally[0].cpu().numpy().copy(), eng.cut_mask[0].cpu().numpy().view(np.uint32).copy(),
                int(eng.cut_count[0].item()))

    def geometry(self):
        """Flat copies of the geometric attributes the compactness updaters read
        (updaters/compactness.py): per node ``boundary_node`` (bool) and ``boundary_perim``
        (0 where absent), per edge of ``self.edges`` ``shared_perim``; plus whether every value
        met was a Python int (the reference's ``sum`` then yields ints)."""
        if self._geometry is None:
            g = self.graph
            on_boundary = np.fromiter((bool(g.nodes[lab].get("boundary_node", False)) for lab in self.labels),
                                      dtype=bool, count=self.n_nodes)
            seen = []
            perim = np.zeros(self.n_nodes, dtype=np.float64)
            for i in np.nonzero(on_boundary)[0].tolist():
                value = g.nodes[self.labels[i]]["boundary_perim"]  # KeyError as in the reference
                perim[i] = value
                seen.append(value)
            shared = np.empty(len(self.edges), dtype=np.float64)
            for e, (u, v, data) in enumerate(g.edges(data=True)):
                value = data["shared_perim"]
                shared[e] = value
                seen.append(value)
            integral = all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in seen)
            self._geometry = (on_boundary, perim, shared, integral)
        return self._geometry

    def region_codes(self, reg_attr: str) -> Tuple[np.ndarray, int]:
        """Dense int32 code per node for the values of a node attribute.  ``None`` is a value
        like any other, as it is for the reference's ``==`` (updaters/county_splits.py:152-156)."""
        hit = self._region_codes.get(reg_attr)
        if hit is None:
            g = self.graph
            codes: Dict = {}
            col = np.fromiter((codes.setdefault(g.nodes[lab][reg_attr], len(codes)) for lab in self.labels),
                              dtype=np.int32, count=self.n_nodes)
            hit = (col, len(codes))
            self._region_codes[reg_attr] = hit
        return hit

    def region_members(self, reg_attr: str):
        """(codes per node, values by code, node labels by code) of a node attribute: the
        per-county node lists of ``county_splits`` (updaters/county_splits.py:75-90), shared by
        every state of the graph."""
        hit = self._region_members.get(reg_attr)
        if hit is None:
            col, n = self.region_codes(reg_attr)
            g = self.graph
            values = [None] * n
            members = [[] for _ in range(n)]
            for lab, code in zip(self.labels, col.tolist()):
                values[code] = g.nodes[lab][reg_attr]
                members[code].append(lab)
            hit = (col, values, members)
            self._region_members[reg_attr] = hit
        return hit

    def region_stats(self, vec: np.ndarray, reg_attr: str, k: int) -> Tuple[int, int, int]:
        col, n_regions = self.region_codes(reg_attr)
        eng = self.engine(("scan",), None, k)
        eng.set_assignment(vec)
        out = eng.region_splits(col, n_regions)[0].cpu().numpy()
        return int(out[0]), int(out[1]), int(out[2])

    def district_pieces(self, vec: np.ndarray, k: int) -> np.ndarray:
        eng = self.engine(("scan",), None, k)
        eng.set_assignment(vec)
        return eng.contiguity()[0].cpu().numpy().copy()
