"""``Grid``: a partition of an m x n lattice with unit populations - the synthetic input of
BASELINE.json configs[0] and [1].

API of the reference's ``Grid`` (/root/reference/gerrychain/grid.py:32-288): nodes are
``(i, j)`` tuples with ``population = 1`` / ``area = 1``, edges carry ``shared_perim``, the
default plan is the four quadrants, ``as_list_of_lists()`` / ``str()`` render the plan.  The
lattice itself comes from ``gerrychain_b200.synthetic.grid_graph``; rendering reads the dense
district vector the engine works on instead of walking a dict.  The default updaters are the
reference's: ``population`` / ``area`` / ``cut_edges`` come from the device, the geometric
ones (perimeter, boundaries, ``polsby_popper``) are host statistics over the same vector.
"""

from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np

from . import synthetic
from .partition import Partition
from .metrics import polsby_popper
from .updaters import (Tally, boundary_nodes, cut_edges, cut_edges_by_part, exterior_boundaries,
                       interior_boundaries, perimeter)


def create_grid_graph(dimensions: Tuple[int, int], with_diagonals: bool = False):
    """m x n lattice graph with the node / edge attributes of grid.py:142-181."""
    try:
        m, n = dimensions
    except (TypeError, ValueError):
        raise ValueError("Expected two dimensions.") from None
    return synthetic.grid_graph(int(m), int(n), with_diagonals)


def color_quadrants(node: Tuple[int, int], thresholds: Tuple[int, int]) -> int:
    """Quadrant index of a lattice node: bit 0 = right of the column threshold... in the
    reference's convention 0/1 by the first coordinate, +2 by the second (grid.py:267-288)."""
    return int(node[0] >= thresholds[0]) + 2 * int(node[1] >= thresholds[1])


class Grid(Partition):
    default_updaters = {  # grid.py:44-54
        "cut_edges": cut_edges,
        "population": Tally("population"),
        "perimeter": perimeter,
        "exterior_boundaries": exterior_boundaries,
        "interior_boundaries": interior_boundaries,
        "boundary_nodes": boundary_nodes,
        "area": Tally("area", alias="area"),
        "polsby_popper": polsby_popper,
        "cut_edges_by_part": cut_edges_by_part,
    }

    def __init__(self, dimensions: Optional[Tuple[int, int]] = None, with_diagonals: bool = False,
                 assignment: Optional[Dict] = None, updaters: Optional[Dict[str, Callable]] = None,
                 parent: Optional["Grid"] = None, flips: Optional[Dict] = None) -> None:
        if parent is not None and not dimensions:
            Partition.__init__(self, parent=parent, flips=flips)  # dimensions travel with _inherit
            return
        if not dimensions:
            raise Exception("Not a good way to create a Partition")
        self.dimensions = tuple(dimensions)
        lattice = create_grid_graph(self.dimensions, with_diagonals)
        if not assignment:
            half = (self.dimensions[0] // 2, self.dimensions[1] // 2)
            assignment = {cell: color_quadrants(cell, half) for cell in lattice.nodes}
        merged = dict(updaters or {})
        merged.update(self.default_updaters)  # the defaults win, as in the reference
        Partition.__init__(self, lattice, assignment, merged)

    def as_list_of_lists(self):
        """Row j of the result lists the districts of cells (0, j) .. (m-1, j)."""
        m, n = self.dimensions
        labels = np.array(self._district_labels, dtype=object)
        cells = labels[self._vec.reshape(m, n)]  # node (i, j) has index i * n + j
        return [list(cells[:, j]) for j in range(n)]

    def __str__(self):
        return "".join("".join(str(d) for d in row) + "\n" for row in self.as_list_of_lists())

    def __repr__(self):
        k = len(self)
        return "{}x{} Grid\nPartitioned into {} part{}".format(self.dimensions[0], self.dimensions[1], k,
                                                              "" if k == 1 else "s")
