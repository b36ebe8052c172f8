"""Updaters of the ReCom path: ``Tally`` and ``cut_edges``.

Reference: /root/reference/gerrychain/updaters/tally.py:69-170 and
updaters/cut_edges.py:99-120.  The reference recomputes both incrementally on the host
from parent/child flows; here both are maintained by the chain-step kernel (per-district
sums and a cut-edge bitmask resident in HBM) and these callables only present the device
values in the reference's shapes (``{part: sum}`` and ``{(u, v), ...}``).
"""

from __future__ import annotations

from typing import List, Optional, Type, Union


class Tally:
    """``Tally(fields, alias=None, dtype=int)`` - per-district sum of node attribute(s)."""

    __slots__ = ["fields", "alias", "dtype"]

    def __init__(self, fields: Union[str, List[str]], alias: Optional[str] = None, dtype: Type = int):
        if not isinstance(fields, list):
            fields = [fields]
        if not alias:
            alias = fields[0]
        self.fields = fields
        self.alias = alias
        self.dtype = dtype

    def __call__(self, partition):
        sums = partition._tally(tuple(self.fields))
        return {part: self.dtype(v) for part, v in zip(partition._district_labels, sums)}


def cut_edges(partition):
    """Set of edges crossing districts, each as ``tuple(sorted((u, v)))``
    (updaters/cut_edges.py:99-120)."""
    return partition._cut_edge_set()


def cut_edges_by_part(partition):
    """``{part: set of cut edges touching it}`` (updaters/cut_edges.py:76-96)."""
    out = {part: set() for part in partition.parts}
    mapping = partition.assignment.mapping
    for e in partition["cut_edges"]:
        out[mapping[e[0]]].add(e)
        out[mapping[e[1]]].add(e)
    return out
