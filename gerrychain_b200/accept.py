"""Acceptance functions (reference: /root/reference/gerrychain/accept.py)."""

import random


def always_accept(partition) -> bool:
    """accept.py:15-16."""
    return True


def cut_edge_accept(partition) -> bool:
    """Metropolis rule on the number of cut edges (accept.py:19-35).  Needs only the
    cut-edge COUNT, which the engine keeps per chain."""
    bound = 1.0
    if partition.parent is not None:
        bound = min(1, partition.parent.n_cut_edges / partition.n_cut_edges)
    return random.random() < bound
