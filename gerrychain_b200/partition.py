"""``Partition`` / ``GeographicPartition`` / ``Assignment`` with GerryChain's surface.

Reference: /root/reference/gerrychain/partition/partition.py:12-213 (Partition),
partition/assignment.py:9-191 (Assignment), partition/geographic.py:13 (GeographicPartition).

What differs underneath: the state of a partition is a dense ``uint8`` (``uint16`` beyond 255
districts) district-index vector in node order (the layout the engine keeps in HBM), not a dict copied at every step
(assignment.py:69-74).  ``Tally`` values and the cut-edge set come from the device: from
the chain-step kernel's incremental updates when the partition was produced by ``recom``,
from ``rc_engine_init_state`` (the first-time scans, updaters/tally.py:118-141 and
updaters/cut_edges.py:110-114) otherwise.  The dict / frozenset views the reference
exposes (``assignment.mapping``, ``parts``, ``flips``) are built lazily on first use.
"""

from __future__ import annotations

from collections.abc import Mapping
from typing import Any, Dict, Optional, Tuple

import numpy as np

from ._context import GraphContext, context_of
from .updaters import (Tally, boundary_nodes, cut_edges, cut_edges_by_part, exterior_boundaries,
                       interior_boundaries, perimeter)


class Assignment(Mapping):
    """node -> part mapping with ``parts`` (part -> frozenset of nodes), built from the
    dense vector on demand (assignment.py:9-132)."""

    __slots__ = ["parts", "mapping"]

    def __init__(self, parts: Dict, mapping: Optional[Dict] = None):
        self.parts = parts
        if mapping is None:
            mapping = {}
            for part, nodes in parts.items():
                for node in nodes:
                    mapping[node] = part
        self.mapping = mapping

    @classmethod
    def _from_vector(cls, labels, vec, district_labels):
        mapping = {lab: district_labels[d] for lab, d in zip(labels, vec.tolist())}
        parts = {part: set() for part in district_labels}
        for lab, part in mapping.items():
            parts[part].add(lab)
        return cls({part: frozenset(nodes) for part, nodes in parts.items()}, mapping)

    def __repr__(self):
        return "<Assignment [{} keys, {} parts]>".format(len(self), len(self.parts))

    def __iter__(self):
        return iter(self.mapping)

    def __len__(self):
        return len(self.mapping)

    def __getitem__(self, node):
        return self.mapping[node]

    def items(self):
        return self.mapping.items()

    def keys(self):
        return self.mapping.keys()

    def values(self):
        return self.mapping.values()

    def copy(self):
        return Assignment(self.parts.copy(), self.mapping.copy())

    def to_dict(self) -> Dict:
        return self.mapping

    def to_series(self):
        import pandas

        groups = [pandas.Series(data=part, index=list(nodes)) for part, nodes in self.parts.items()]
        return pandas.concat(groups)

    @classmethod
    def from_dict(cls, assignment: Dict) -> "Assignment":
        parts: Dict[Any, set] = {}
        for node, part in assignment.items():
            parts.setdefault(part, set()).add(node)
        return cls({part: frozenset(nodes) for part, nodes in parts.items()})


def get_assignment(part_assignment, graph=None) -> Dict:
    """assignment.py:154-191: node-attribute key, dict-like or Assignment -> node -> part."""
    if isinstance(part_assignment, str):
        if graph is None:
            raise TypeError("You must provide a graph when using a node attribute for the part_assignment")
        return {node: graph.nodes[node][part_assignment] for node in graph.nodes}
    if isinstance(part_assignment, Assignment):
        return part_assignment.mapping
    if callable(getattr(part_assignment, "items", None)):
        return dict(part_assignment.items())
    raise TypeError("Assignment must be a dict or a node attribute key")


_CORE = ("graph", "updaters", "parent", "_flips", "_cache", "_ctx", "_vec", "_district_labels",
         "_rank", "_assignment", "_dev_tally", "_cut_mask", "_cut_count", "_pair")


class _SubgraphView:
    """Lazy ``part -> graph.subgraph(nodes of part)`` (partition/subgraphs.py:5-64)."""

    def __init__(self, graph, parts):
        self.graph, self.parts, self._built = graph, parts, {}

    def __getitem__(self, part):
        if part not in self._built:
            self._built[part] = self.graph.subgraph(self.parts[part])
        return self._built[part]

    def __iter__(self):
        return (self[part] for part in self.parts)

    def items(self):
        return ((part, self[part]) for part in self.parts)

    def __repr__(self):
        return f"<SubgraphView with {len(self.parts)} and {len(self._built)} cached graphs>"


class Partition:
    """Partition of the nodes of a graph into districts (partition.py:12)."""

    default_updaters = {"cut_edges": cut_edges}

    def __init__(self, graph=None, assignment=None, updaters=None, parent=None, flips=None,
                 use_default_updaters=True):
        if parent is None:
            self._first_time(graph, assignment, updaters, use_default_updaters)
        else:
            self._from_parent(parent, flips)

    # -- construction ------------------------------------------------------------------
    def _first_time(self, graph, assignment, updaters, use_default_updaters):
        ctx = context_of(graph)  # TypeError for anything that is not a graph
        self._ctx: GraphContext = ctx
        self.graph = graph
        mapping = get_assignment(assignment, ctx.graph)
        if set(mapping) != set(ctx.labels):
            raise KeyError("The graph's node labels do not match the Assignment's keys")
        labels = sorted(set(mapping.values()))  # tree_proposals.py:113 relies on this order
        if len(labels) > 32767:
            raise ValueError("the device engine stores district ids in 16 bits: at most 32767 districts")
        self._district_labels = labels
        self._rank = {p: i for i, p in enumerate(labels)}
        # one byte per node up to 255 districts, two beyond (the engine's two label widths)
        self._vec = np.fromiter((self._rank[mapping[lab]] for lab in ctx.labels),
                                dtype=np.uint8 if len(labels) <= 255 else np.uint16, count=ctx.n_nodes)
        if updaters is None:
            updaters = dict()
        self.updaters = dict(self.default_updaters) if use_default_updaters else {}
        self.updaters.update(updaters)
        self.parent = None
        self._flips = None
        self._reset_caches()

    def _reset_caches(self):
        self._cache = dict()
        self._assignment = None
        self._dev_tally = {}
        self._cut_mask = None
        self._cut_count = None
        self._pair = None

    def _from_parent(self, parent: "Partition", flips: Dict) -> None:
        self._inherit(parent)
        self._flips = dict(flips)
        vec = parent._vec.copy()
        index = self._ctx.index
        try:
            for node, part in flips.items():
                vec[index[node]] = self._rank[part]
        except KeyError as e:
            raise KeyError(f"flip refers to an unknown node or district: {e}") from None
        self._vec = vec

    def _inherit(self, parent):
        for k, v in parent.__dict__.items():  # subclasses' own attributes (e.g. Grid.dimensions)
            if k not in _CORE:
                self.__dict__[k] = v
        self.parent = parent
        self.graph = parent.graph
        self.updaters = parent.updaters
        self._ctx = parent._ctx
        self._district_labels = parent._district_labels
        self._rank = parent._rank
        self._reset_caches()

    @classmethod
    def _from_device(cls, parent: "Partition", vec, tally_fields, tally, cut_count, cut_mask, pair):
        """Child produced by the chain-step kernel (proposals.recom)."""
        self = cls.__new__(cls)
        self._inherit(parent)
        self._flips = None
        # (the 16-bit-label engine hands back signed torch rows; labels <= 32767, so the cast is exact)
        self._vec = np.asarray(vec).astype(parent._vec.dtype, copy=False)
        self._dev_tally[tuple(tally_fields)] = tally
        self._cut_count = int(cut_count)
        self._cut_mask = cut_mask
        self._pair = pair
        return self

    @classmethod
    def from_random_assignment(cls, graph, n_parts, epsilon, pop_col, updaters=None,
                               use_default_updaters=True, flips=None, method=None):
        """partition.py:67-120: a partition from a random seed plan drawn by
        ``recursive_tree_part`` (tree.py:971-1085) - here by ``rc_engine_seed``."""
        from .seed import recursive_tree_part

        ctx = context_of(graph)
        total_pop = int(ctx.column(pop_col).sum())
        ideal_pop = total_pop / n_parts
        if method is None:
            assignment = recursive_tree_part(graph, range(n_parts), ideal_pop, pop_col, epsilon)
        else:
            assignment = method(graph=graph, parts=range(n_parts), pop_target=ideal_pop, pop_col=pop_col,
                                epsilon=epsilon)
        return cls(graph, assignment, updaters, use_default_updaters=use_default_updaters)

    @classmethod
    def from_districtr_file(cls, graph, districtr_file: str, updaters=None):
        """partition.py:253-296: a plan exported from Districtr - ``assignment`` keyed by the
        unit id found in the node attribute named by ``idColumn.key``."""
        import json

        with open(districtr_file) as handle:
            exported = json.load(handle)
        id_key = exported["idColumn"]["key"]
        plan = exported["assignment"]
        try:
            unit_of = {node: str(graph.nodes[node][id_key]) for node in graph}
        except KeyError:
            raise TypeError(f"The provided graph is missing the {id_key} column, which is needed to "
                            "match the Districtr assignment to the nodes of the graph.") from None
        return cls(graph, {node: plan[unit_of[node]] for node in graph}, updaters)

    # -- reference surface -------------------------------------------------------------
    def __repr__(self):
        number_of_parts = len(self)
        s = "s" if number_of_parts > 1 else ""
        return "<{} [{} part{}]>".format(self.__class__.__name__, number_of_parts, s)

    def __len__(self):
        return len(self._district_labels)

    def flip(self, flips: Dict) -> "Partition":
        return self.__class__(parent=self, flips=flips)

    def crosses_parts(self, edge: Tuple) -> bool:
        i = self._ctx.index
        return bool(self._vec[i[edge[0]]] != self._vec[i[edge[1]]])

    def __getitem__(self, key: str) -> Any:
        if key not in self._cache:
            self._cache[key] = self.updaters[key](self)
        return self._cache[key]

    def __getattr__(self, key):
        if key.startswith("_") or key in _CORE:
            raise AttributeError(key)
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def keys(self):
        return self.updaters.keys()

    @property
    def assignment(self) -> Assignment:
        if self._assignment is None:
            self._assignment = Assignment._from_vector(self._ctx.labels, self._vec, self._district_labels)
        return self._assignment

    @property
    def parts(self):
        return self.assignment.parts

    @property
    def subgraphs(self):
        """``{part: induced subgraph}``, built on first use (partition/subgraphs.py)."""
        return _SubgraphView(self.graph, self.parts)

    @property
    def flips(self):
        """Parent -> child changes.  For a ReCom child: every node of the two merged
        districts with its new district (tree.py:947-961), as in the reference."""
        if self._flips is None and self._pair is not None and self._pair[0] >= 0:
            a, b = int(self._pair[0]), int(self._pair[1])
            old = self.parent._vec if self.parent is not None else self._vec
            idx = np.nonzero((old == a) | (old == b))[0]
            labels, dl = self._ctx.labels, self._district_labels
            self._flips = {labels[i]: dl[int(self._vec[i])] for i in idx.tolist()}
        return self._flips

    # -- device-backed values ----------------------------------------------------------
    def _scan(self, fields):
        tally, mask, count = self._ctx.scan(self._vec, fields, len(self._district_labels))
        if fields is not None:
            self._dev_tally[tuple(fields)] = tally
        if self._cut_mask is None:
            self._cut_mask, self._cut_count = mask, count

    def _tally(self, fields: Tuple[str, ...]) -> np.ndarray:
        fields = tuple(fields)
        if fields not in self._dev_tally:
            col = self._ctx.float_column(fields)
            if col is None:
                self._scan(fields)
            else:
                # A column the device cannot tally exactly (floats - e.g. the `area` of
                # GeographicPartition's default updaters - NaNs, or a total beyond int32): summed
                # on the host over the dense plan.  The reference's Tally sums whatever the node
                # attribute holds and skips NaN with a warning (updaters/tally.py:118-141).
                import warnings

                bad = np.isnan(col)
                for i in np.nonzero(bad)[0].tolist():
                    warnings.warn("ignoring nan encountered at node '{}' for attribute '{}'".format(
                        self._ctx.labels[i], fields[0]))
                self._dev_tally[fields] = np.bincount(self._vec[~bad], weights=col[~bad],
                                                      minlength=len(self._district_labels))
        return self._dev_tally[fields]

    def _region_stats(self, reg_attr: str):
        """(regions with a cut edge inside, regions over > 1 district, pieces) of this plan
        for the node attribute ``reg_attr`` (``rc_engine_region_splits``)."""
        key = ("region_stats", reg_attr)
        if key not in self._cache:
            self._cache[key] = self._ctx.region_stats(self._vec, reg_attr, len(self._district_labels))
        return self._cache[key]

    def _district_pieces(self) -> np.ndarray:
        """Connected pieces of every district (``rc_engine_contiguity``)."""
        key = ("district_pieces",)
        if key not in self._cache:
            self._cache[key] = self._ctx.district_pieces(self._vec, len(self._district_labels))
        return self._cache[key]

    @property
    def n_cut_edges(self) -> int:
        """``len(partition["cut_edges"])`` without materialising the set."""
        if self._cut_count is None:
            self._scan(None)
        return self._cut_count

    def _cut_edge_set(self):
        if self._cut_mask is None:
            self._scan(None)
        csr = self._ctx.csr(None)
        bits = np.unpackbits(self._cut_mask.view(np.uint8), bitorder="little")[: csr.n_edges]
        lab = self._ctx.edge_labels(csr)
        return {lab[e] for e in np.nonzero(bits)[0].tolist()}

    def assignment_vector(self) -> np.ndarray:
        """District index per node, node order (read-only view of the engine's layout)."""
        v = self._vec.view()
        v.flags.writeable = False
        return v


class GeographicPartition(Partition):
    """partition/geographic.py:13: a ``Partition`` with areas, perimeters and boundary
    information, the inputs of compactness scores.  The geometric updaters are host
    statistics over the dense district vector (``updaters.perimeter`` etc.); ``area`` is a
    ``Tally`` like any other."""

    default_updaters = {
        "perimeter": perimeter,
        "exterior_boundaries": exterior_boundaries,
        "interior_boundaries": interior_boundaries,
        "boundary_nodes": boundary_nodes,
        "cut_edges": cut_edges,
        "area": Tally("area", alias="area"),
        "cut_edges_by_part": cut_edges_by_part,
    }
