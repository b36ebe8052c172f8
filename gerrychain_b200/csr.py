"""Host-side graph layout for the device engine: canonical CSR + edge ids.

The reference keeps the dual graph as a networkx dict-of-dicts
(/root/reference/gerrychain/graph/graph.py:51, FrozenGraph :508).  The engine
wants flat arrays.  Conventions (DESIGN.md "Data layout"):

* node index  = position of the node in ``graph.nodes`` iteration order;
* undirected edge id = index in the lexicographically sorted list of
  ``(min(u, v), max(u, v))`` node-index pairs;
* CSR rows are sorted by neighbour index, so enumerating ``(u, v), u < v`` row by
  row visits edges in ascending edge id;
* ``slot_edge`` gives the undirected edge id of every directed CSR slot.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Hashable, List, Mapping, Optional, Sequence

import numpy as np


@dataclass
class GraphCSR:
    """Flat description of the dual graph (host copy, numpy)."""

    n_nodes: int
    n_edges: int
    row_ptr: np.ndarray  # int32 [N+1]
    col_idx: np.ndarray  # int32 [2E]
    slot_edge: np.ndarray  # int32 [2E]   undirected edge id per directed slot
    edge_uv: np.ndarray  # int32 [E, 2]  (u < v), sorted lexicographically
    pop: np.ndarray  # int32 [N]
    region_ids: np.ndarray = field(
        default_factory=lambda: np.zeros((0, 0), dtype=np.int32)
    )  # int32 [R, N]; -1 == None in the reference
    surcharges: np.ndarray = field(
        default_factory=lambda: np.zeros((0,), dtype=np.float64)
    )  # float64 [R]
    node_labels: Optional[List[Hashable]] = None  # index -> reference node id
    region_names: Sequence[str] = ()

    @property
    def n_region_cols(self) -> int:
        return int(self.region_ids.shape[0])

    def node_index(self) -> Dict[Hashable, int]:
        labels = self.node_labels
        if labels is None:
            return {i: i for i in range(self.n_nodes)}
        return {lab: i for i, lab in enumerate(labels)}

    def edge_id_lookup(self) -> Dict[tuple, int]:
        return {(int(u), int(v)): i for i, (u, v) in enumerate(self.edge_uv)}

    def edge_id(self, u: int, v: int) -> int:
        """Edge id of the undirected edge {u, v} given node indices."""
        if u > v:
            u, v = v, u
        lo, hi = int(self.row_ptr[u]), int(self.row_ptr[u + 1])
        j = lo + int(np.searchsorted(self.col_idx[lo:hi], v))
        if j >= hi or self.col_idx[j] != v:
            raise KeyError((u, v))
        return int(self.slot_edge[j])

    def with_regions(
        self, region_ids: np.ndarray, surcharges: Sequence[float], names=()
    ) -> "GraphCSR":
        region_ids = np.ascontiguousarray(region_ids, dtype=np.int32).reshape(
            -1, self.n_nodes
        )
        sur = np.asarray(surcharges, dtype=np.float64).reshape(-1)
        assert region_ids.shape[0] == sur.shape[0]
        return GraphCSR(
            self.n_nodes,
            self.n_edges,
            self.row_ptr,
            self.col_idx,
            self.slot_edge,
            self.edge_uv,
            self.pop,
            region_ids,
            sur,
            self.node_labels,
            tuple(names),
        )


def csr_from_edges(
    n_nodes: int,
    edges: np.ndarray,
    pop: np.ndarray,
    node_labels: Optional[List[Hashable]] = None,
) -> GraphCSR:
    """Build the canonical CSR from an (E, 2) array of node-index pairs."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    lo = np.minimum(edges[:, 0], edges[:, 1])
    hi = np.maximum(edges[:, 0], edges[:, 1])
    if np.any(lo == hi):
        raise ValueError("self loops are not supported")
    uv = np.unique(np.stack([lo, hi], axis=1), axis=0)  # sorted lexicographically
    n_edges = uv.shape[0]
    eid = np.arange(n_edges, dtype=np.int64)
    src = np.concatenate([uv[:, 0], uv[:, 1]])
    dst = np.concatenate([uv[:, 1], uv[:, 0]])
    sid = np.concatenate([eid, eid])
    order = np.lexsort((dst, src))
    src, dst, sid = src[order], dst[order], sid[order]
    row_ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    pop = np.asarray(pop)
    if pop.shape != (n_nodes,):
        raise ValueError("population vector has the wrong length")
    if np.any(pop != np.floor(pop)):
        raise TypeError("the engine tallies integer populations only")
    if np.abs(pop).sum() >= 2**31:
        raise OverflowError("total population must fit in int32")
    return GraphCSR(
        n_nodes=int(n_nodes),
        n_edges=int(n_edges),
        row_ptr=row_ptr.astype(np.int32),
        col_idx=dst.astype(np.int32),
        slot_edge=sid.astype(np.int32),
        edge_uv=uv.astype(np.int32),
        pop=pop.astype(np.int32),
        node_labels=node_labels,
    )


def csr_from_networkx(
    graph: Any,
    pop_col: str,
    region_surcharge: Optional[Mapping[str, float]] = None,
) -> GraphCSR:
    """CSR of a networkx-style graph (the reference's ``Graph`` / ``FrozenGraph``).

    ``region_surcharge`` mirrors ``recom(..., region_surcharge=)``
    (/root/reference/gerrychain/proposals/tree_proposals.py:42): one region column
    per key, in dict order (the order the reference adds surcharges in,
    /root/reference/gerrychain/tree.py:80-87).  Region labels may be any hashable;
    ``None`` maps to -1 (always surcharged, tree.py:84-85).
    """
    g = graph  # FrozenGraph forwards .nodes / .edges to the wrapped Graph
    labels = list(g.nodes)
    index = {lab: i for i, lab in enumerate(labels)}
    edges = np.array([(index[u], index[v]) for u, v in g.edges], dtype=np.int64)
    pop = np.array([g.nodes[lab][pop_col] for lab in labels])
    csr = csr_from_edges(len(labels), edges, pop, node_labels=labels)
    if region_surcharge:
        cols = []
        for key in region_surcharge:
            codes: Dict[Any, int] = {}
            col = np.empty(len(labels), dtype=np.int32)
            for i, lab in enumerate(labels):
                val = g.nodes[lab].get(key)
                col[i] = -1 if val is None else codes.setdefault(val, len(codes))
            cols.append(col)
        csr = csr.with_regions(
            np.stack(cols), list(region_surcharge.values()), tuple(region_surcharge)
        )
    return csr


def assignment_vector(csr: GraphCSR, assignment: Mapping[Hashable, Any]):
    """Dense district index per node plus the sorted list of district labels.

    District index = rank of the label in ``sorted(labels)``, which preserves the
    ordering the reference relies on when it sorts the merged pair
    (/root/reference/gerrychain/proposals/tree_proposals.py:113).
    """
    labels = csr.node_labels if csr.node_labels is not None else range(csr.n_nodes)
    parts = sorted(set(assignment[lab] for lab in labels))
    rank = {p: i for i, p in enumerate(parts)}
    vec = np.fromiter(
        (rank[assignment[lab]] for lab in labels), dtype=np.int32, count=csr.n_nodes
    )
    return vec, parts
