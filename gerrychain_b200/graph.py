"""``Graph``: the dual graph as input format, with GerryChain's JSON I/O.

Reference: /root/reference/gerrychain/graph/graph.py:51-135 (``Graph``, ``from_networkx``,
``from_json``, ``to_json``, ``node_indices``, ``lookup``).  The on-disk format is networkx's
``json_graph`` adjacency layout (``nodes[]`` with attributes + ``id``, ``adjacency[][]``), the
wire format GerryChain users keep their dual graphs in.
Shapefile / GeoDataFrame constructors are GIS preprocessing and out of scope; the engine
itself takes any networkx-style graph.
"""

from __future__ import annotations

import json
import warnings
from typing import Any

import networkx
from networkx.readwrite import json_graph


class Graph(networkx.Graph):
    def __repr__(self):
        return "<Graph [{} nodes, {} edges]>".format(len(self.nodes), len(self.edges))

    @classmethod
    def from_networkx(cls, graph: networkx.Graph) -> "Graph":
        return cls(graph)

    @classmethod
    def from_json(cls, json_file: str) -> "Graph":
        with open(json_file) as f:
            data = json.load(f)
        graph = cls.from_networkx(json_graph.adjacency_graph(data))
        graph.issue_warnings()
        return graph

    def to_json(self, json_file: str, *, include_geometries_as_geojson: bool = False) -> None:
        data = json_graph.adjacency_data(self)
        for node in data["nodes"]:  # geometry objects are not serialisable (graph.py:106-109)
            node.pop("geometry", None)
        with open(json_file, "w") as f:
            json.dump(data, f, default=lambda x: x.item() if hasattr(x, "item") else str(x))

    @property
    def node_indices(self):
        return set(self.nodes)

    @property
    def edge_indices(self):
        return set(self.edges)

    def lookup(self, node: Any, field: Any) -> Any:
        return self.nodes[node][field]

    @property
    def islands(self):
        return set(node for node in self if self.degree[node] == 0)

    def warn_for_islands(self) -> None:
        islands = self.islands
        if len(self.islands) > 0:
            warnings.warn("Found islands (degree-0 nodes). Indices of islands: {}".format(islands))

    def issue_warnings(self) -> None:
        self.warn_for_islands()
