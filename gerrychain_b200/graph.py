"""``Graph``: the dual graph as input format, with GerryChain's JSON wire format.

Mirrors the part of /root/reference/gerrychain/graph/graph.py:51-135 that feeds the hot path:
a ``networkx.Graph`` subclass with ``from_networkx``, ``from_json`` / ``to_json`` (networkx
``json_graph`` adjacency layout: ``nodes[]`` with attributes and ``id``, ``adjacency[][]``),
``node_indices``, ``lookup`` and the island warning.  Shapefile / GeoDataFrame constructors
are GIS preprocessing and out of scope; the engine itself accepts any
networkx-style graph, this class only adds the reference's I/O (section 8f rank 4).
"""

from __future__ import annotations

import json
import warnings

import networkx
from networkx.readwrite.json_graph import adjacency_data, adjacency_graph


def _jsonable(value):
    """numpy scalars -> Python numbers, anything else unserialisable -> its string."""
    item = getattr(value, "item", None)
    return item() if callable(item) else str(value)


class Graph(networkx.Graph):
    def __repr__(self):
        return f"<Graph [{self.number_of_nodes()} nodes, {self.number_of_edges()} edges]>"

    # -- constructors -------------------------------------------------------------------
    @classmethod
    def from_networkx(cls, graph: networkx.Graph) -> "Graph":
        return cls(graph)

    @classmethod
    def from_json(cls, json_file: str) -> "Graph":
        with open(json_file) as handle:
            loaded = cls(adjacency_graph(json.load(handle)))
        loaded.issue_warnings()
        return loaded

    def to_json(self, json_file: str, *, include_geometries_as_geojson: bool = False) -> None:
        payload = adjacency_data(self)
        if not include_geometries_as_geojson:
            for record in payload["nodes"]:
                record.pop("geometry", None)  # shapely objects cannot be serialised
        with open(json_file, "w") as handle:
            json.dump(payload, handle, default=_jsonable)

    # -- lookups ------------------------------------------------------------------------
    @property
    def node_indices(self):
        return set(self.nodes)

    @property
    def edge_indices(self):
        return set(self.edges)

    def lookup(self, node, field):
        return self.nodes[node][field]

    @property
    def islands(self):
        return {v for v, degree in self.degree if degree == 0}

    def issue_warnings(self) -> None:
        lonely = self.islands
        if lonely:
            warnings.warn(f"Found islands (degree-0 nodes). Indices of islands: {lonely}")
