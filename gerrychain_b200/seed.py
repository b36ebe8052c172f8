"""Seed plans on the device: ``recursive_tree_part`` / ``Partition.from_random_assignment``.

Reference: /root/reference/gerrychain/tree.py:971-1085 and partition/partition.py:67-120.
One CTA builds one plan (``rc_engine_seed``); ``random_plans`` draws many at once, e.g. one
per chain of a ``BatchedMarkovChain``.
"""

from __future__ import annotations

import random
from typing import Optional

import numpy as np

from ._context import context_of
from .config import EngineConfig


def random_plans(graph, n_parts: int, epsilon: float, pop_col: str, n_plans: int = 1,
                 pop_target: Optional[float] = None, seed: Optional[int] = None,
                 max_attempts: Optional[int] = 10000, device=None) -> np.ndarray:
    """``n_plans`` independent random plans as a ``uint8 [n_plans, N]`` array (``uint16``
    beyond 255 districts) of district indices (node order of ``graph.nodes``); every district within ``epsilon`` of
    ``pop_target`` (default: total / n_parts) and contiguous."""
    from .engine import Engine
    from .proposals import raise_for_status

    ctx = context_of(graph)
    csr = ctx.csr(pop_col)
    if pop_target is None:
        pop_target = float(csr.pop.sum()) / n_parts
    cfg = EngineConfig(n_districts=int(n_parts), n_chains=int(n_plans), pop_target=float(pop_target),
                       epsilon=float(epsilon), seed=random.getrandbits(64) if seed is None else int(seed),
                       cap_nodes=csr.n_nodes, cap_edges=csr.n_edges)
    eng = Engine(csr, cfg, np.zeros(csr.n_nodes, dtype=np.uint8), device=device)
    eng.seed(max_attempts).synchronize()
    status = eng.status.cpu().numpy()
    bad = np.nonzero(status)[0]
    if len(bad):
        raise_for_status(int(status[bad[0]]), max_attempts)
    plans = eng.assignments().copy()
    eng.close()
    return plans


def recursive_tree_part(graph, parts, pop_target, pop_col, epsilon, node_repeats=1, method=None):
    """tree.py:971-1085: ``{node: part}`` for ``len(parts)`` balanced, contiguous parts."""
    from .proposals import method_config

    if node_repeats != 1:
        raise NotImplementedError("the device path draws a new tree for every attempt (node_repeats=1)")
    max_attempts = 10000  # the reference's default: partial(bipartition_tree, max_attempts=10000)
    if method is not None:
        mc = method_config(method)
        if mc["tree_kind"] != 0:
            raise NotImplementedError("seed plans use random_spanning_tree")
        max_attempts = mc["max_attempts"]
    parts = list(parts)
    plan = random_plans(graph, len(parts), epsilon, pop_col, 1, pop_target=pop_target,
                        max_attempts=max_attempts)[0]
    labels = context_of(graph).labels
    return {lab: parts[int(d)] for lab, d in zip(labels, plan.tolist())}
