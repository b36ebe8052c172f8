"""Engine configuration: the kwargs of ``recom`` / ``bipartition_tree`` / the population
``Bounds`` flattened into the plain C struct ``RcConfig`` (include/recom_b200.h).

Reference call sites: /root/reference/gerrychain/proposals/tree_proposals.py:36-44,
/root/reference/gerrychain/tree.py:581-597, constraints/validity.py:63-93.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from . import _abi


@dataclass
class EngineConfig:
    n_districts: int
    n_chains: int
    pop_target: float
    epsilon: float
    node_repeats: int = 1
    max_attempts: Optional[int] = 100000  # bipartition_tree default, tree.py:593
    warn_attempts: int = 1000  # tree.py:594
    allow_pair_reselection: bool = False
    tree_kind: int = _abi.RC_TREE_MST
    cut_policy: int = _abi.RC_CUT_UNIFORM
    pop_lower: float = 1.0  # lower > upper: no Bounds check
    pop_upper: float = 0.0
    seed: int = 0
    chain_offset: int = 0
    cap_nodes: int = 0
    cap_edges: int = 0
    threads_per_cta: int = 0
    max_invalid_streak: int = 0
    accept_rule: int = _abi.RC_ACCEPT_ALWAYS
    proposal_kind: int = _abi.RC_PROPOSAL_RECOM
    rev_M: int = 1
    wilson_walkers: int = 0  # Wilson: trees of one bipartition_tree call drawn concurrently, one walk per lane (0 = engine picks)

    def to_struct(self) -> _abi.RcConfig:
        c = _abi.RcConfig()
        c.n_districts = int(self.n_districts)
        c.n_chains = int(self.n_chains)
        c.pop_target = float(self.pop_target)
        c.epsilon = float(self.epsilon)
        c.node_repeats = int(self.node_repeats)
        c.max_attempts = int(self.max_attempts) if self.max_attempts else 0
        c.warn_attempts = int(self.warn_attempts)
        c.allow_pair_reselection = int(bool(self.allow_pair_reselection))
        c.tree_kind = int(self.tree_kind)
        c.cut_policy = int(self.cut_policy)
        c.pop_lower = float(self.pop_lower)
        c.pop_upper = float(self.pop_upper)
        c.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        c.chain_offset = int(self.chain_offset)
        c.cap_nodes = int(self.cap_nodes)
        c.cap_edges = int(self.cap_edges)
        c.threads_per_cta = int(self.threads_per_cta)
        c.max_invalid_streak = int(self.max_invalid_streak)
        c.accept_rule = int(self.accept_rule)
        c.proposal_kind = int(self.proposal_kind)
        c.rev_M = int(self.rev_M)
        c.wilson_walkers = int(self.wilson_walkers)
        return c


def cut_policy_for(region_surcharge, tree_kind: int) -> int:
    """Which branch of ``_region_preferred_max_weight_choice`` (tree.py:501-578) runs."""
    if tree_kind == _abi.RC_TREE_WILSON or not isinstance(region_surcharge, dict):
        return _abi.RC_CUT_UNIFORM
    if 1 <= len(region_surcharge) <= 3:
        return _abi.RC_CUT_REGION_PREF
    return _abi.RC_CUT_MAX_WEIGHT


def strides(n_nodes: int, n_edges: int):
    """(assign_stride labels, mask_stride words): 16-byte aligned rows (128-bit loads of the assignment row)."""
    return (n_nodes + 15) // 16 * 16, ((n_edges + 31) // 32 + 3) // 4 * 4
