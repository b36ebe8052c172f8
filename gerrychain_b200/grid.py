"""``Grid``: a partition of an m x n lattice with unit populations.

Reference: /root/reference/gerrychain/grid.py:32-288 - same nodes ``(i, j)``, attributes
(``population = 1``, ``area = 1``, ``shared_perim``), default quadrant plan
(``color_quadrants``) and ``as_list_of_lists``.  It is the synthetic input of BASELINE.json
configs[0] and [1].  The geometric default updaters of the reference's Grid (perimeter,
boundaries, polsby_popper) are downstream statistics outside the accelerated path.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

from . import synthetic
from .partition import Partition
from .updaters import Tally, cut_edges, cut_edges_by_part


def create_grid_graph(dimensions: Tuple[int, int], with_diagonals: bool = False):
    """grid.py:142-181."""
    if len(dimensions) != 2:
        raise ValueError("Expected two dimensions.")
    return synthetic.grid_graph(dimensions[0], dimensions[1], with_diagonals)


def color_quadrants(node: Tuple[int, int], thresholds: Tuple[int, int]) -> int:
    """grid.py:267-288."""
    x, y = node
    x_color = 0 if x < thresholds[0] else 1
    y_color = 0 if y < thresholds[1] else 2
    return x_color + y_color


class Grid(Partition):
    default_updaters = {
        "cut_edges": cut_edges,
        "population": Tally("population"),
        "area": Tally("area", alias="area"),
        "cut_edges_by_part": cut_edges_by_part,
    }

    def __init__(self, dimensions: Optional[Tuple[int, int]] = None, with_diagonals: bool = False,
                 assignment: Optional[Dict] = None, updaters: Optional[Dict[str, Callable]] = None,
                 parent: Optional["Grid"] = None, flips: Optional[Dict] = None) -> None:
        if dimensions:
            self.dimensions = dimensions
            graph = create_grid_graph(dimensions, with_diagonals)
            if not assignment:
                thresholds = tuple(math.floor(n / 2) for n in self.dimensions)
                assignment = {node: color_quadrants(node, thresholds) for node in graph.nodes}
            if not updaters:
                updaters = dict()
            updaters.update(self.default_updaters)
            super().__init__(graph, assignment, updaters)
        elif parent:
            super().__init__(parent=parent, flips=flips)
        else:
            raise Exception("Not a good way to create a Partition")

    def __str__(self):
        rows = self.as_list_of_lists()
        return "\n".join(["".join([str(x) for x in row]) for row in rows]) + "\n"

    def __repr__(self):
        dims = "x".join(str(d) for d in self.dimensions)
        number_of_parts = len(self.parts)
        s = "s" if number_of_parts > 1 else ""
        return "{} Grid\nPartitioned into {} part{}".format(dims, number_of_parts, s)

    def as_list_of_lists(self):
        m, n = self.dimensions
        return [[self.assignment.mapping[(i, j)] for i in range(m)] for j in range(n)]
