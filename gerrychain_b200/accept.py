"""Acceptance rules of a chain step (the third argument of ``MarkovChain``).

Reference: /root/reference/gerrychain/accept.py.  ``always_accept`` is the rule of the ReCom
path and the one the batched device runner implements; ``cut_edge_accept`` is the reference's
Metropolis rule on the number of cut edges, evaluated here from the per-state cut-edge COUNT
the engine maintains instead of from two materialised edge sets.
"""

import random


def always_accept(partition) -> bool:
    """Every valid proposal is taken (accept.py:15-16)."""
    return True


def cut_edge_accept(partition) -> bool:
    """More cut edges: accept; fewer: accept with probability old/new... i.e. the ratio
    ``len(parent cut_edges) / len(cut_edges)`` capped at 1 (accept.py:19-35)."""
    if partition.parent is None:
        return True if random.random() < 1.0 else False
    ratio = partition.parent.n_cut_edges / partition.n_cut_edges
    return random.random() < (ratio if ratio < 1 else 1.0)
