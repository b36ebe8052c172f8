"""``recom`` / ``ReCom`` with GerryChain's signature, executed by the CUDA engine.

Reference: /root/reference/gerrychain/proposals/tree_proposals.py:36-146 (recom), :270-305
(ReCom), :18 (MetagraphError).  ``recom(partition, ...)`` returns a child ``Partition``;
the whole proposal - pick a cut edge, merge the two districts, draw spanning trees until
one has a balanced edge, cut, relabel, update Tally and cut_edges - is ONE launch of the
chain-step kernel through ``rc_engine_step_host`` (host plan in, host plan out), so an
unmodified GerryChain driver loop (``MarkovChain``) works on top of it.  For throughput
use ``gerrychain_b200.BatchedMarkovChain``, which keeps thousands of chains resident.

Randomness: the reference draws from Python's global ``random``; the engine draws from a
Philox stream whose key is taken from ``random.getrandbits(64)`` when the engine for a
(graph, configuration) is first built - ``random.seed(s)`` before the first proposal makes
a run reproducible, as in the reference.
"""

from __future__ import annotations

import functools
import warnings
from typing import Callable, Dict, Optional, Union

import numpy as np

from . import _abi, tree
from .config import cut_policy_for
from .tree import BipartitionWarning, bipartition_tree


class MetagraphError(Exception):
    """tree_proposals.py:18-24."""


class ValueWarning(UserWarning):
    """tree_proposals.py:27-33."""


class ReversibilityError(Exception):
    """tree_proposals.py:308-312."""

    def __init__(self, msg):
        self.message = msg


def method_config(method) -> Dict:
    """Engine settings selected by ``method`` (``bipartition_tree`` or a ``partial`` of it,
    tree.py:581-597).  Hooks the engine does not implement raise instead of being ignored."""
    kw = dict(tree.ENGINE_KWARGS)
    func, layers = method, []
    while isinstance(func, functools.partial):
        if func.args:
            raise NotImplementedError("positional arguments bound to bipartition_tree are not supported")
        layers.append(func.keywords)
        func = func.func
    for kws in reversed(layers):  # the outermost partial wins
        kw.update(kws)
    if func is not bipartition_tree:
        raise NotImplementedError(
            f"recom(method={getattr(func, '__name__', func)!r}): the CUDA engine implements "
            "bipartition_tree (tree.py:581); other bipartition methods have no device path")
    import random

    if kw.get("choice", random.choice) is random.choice:  # the reference's default, spelled out
        kw.pop("choice", None)
    unknown = set(kw) - set(tree.ENGINE_KWARGS) - {"region_surcharge"}
    if unknown:
        raise NotImplementedError(f"bipartition_tree hooks without a device path: {sorted(unknown)}")
    if kw["spanning_tree_fn"] is tree.random_spanning_tree:
        tree_kind = _abi.RC_TREE_MST
    elif kw["spanning_tree_fn"] is tree.uniform_spanning_tree:
        tree_kind = _abi.RC_TREE_WILSON
    else:
        raise NotImplementedError("spanning_tree_fn must be random_spanning_tree or uniform_spanning_tree")
    if kw["balance_edge_fn"] not in (tree.find_balanced_edge_cuts_memoization,
                                     tree.find_balanced_edge_cuts_contraction):
        raise NotImplementedError("balance_edge_fn must be find_balanced_edge_cuts_memoization or "
                                  "find_balanced_edge_cuts_contraction (the same set of balanced cuts)")
    if kw["one_sided_cut"]:
        raise NotImplementedError("one_sided_cut=True is the seeding path, not ReCom (tree.py:386-409)")
    if kw["cut_choice"] is tree._region_preferred_max_weight_choice:
        cut = "region"
    elif kw["cut_choice"] is tree._max_weight_choice:
        cut = "max"
    else:
        raise NotImplementedError("cut_choice must be _region_preferred_max_weight_choice or _max_weight_choice")
    return dict(tree_kind=tree_kind, cut=cut, max_attempts=kw["max_attempts"],
                warn_attempts=kw["warn_attempts"], allow_pair_reselection=bool(kw["allow_pair_reselection"]))


def engine_kwargs(pop_target, epsilon, node_repeats, region_surcharge, method) -> Dict:
    """RcConfig fields for a recom configuration (shared by recom and BatchedMarkovChain)."""
    m = method_config(method)
    if m["cut"] == "max":
        policy = _abi.RC_CUT_UNIFORM if m["tree_kind"] == _abi.RC_TREE_WILSON else _abi.RC_CUT_MAX_WEIGHT
    else:
        policy = cut_policy_for(region_surcharge, m["tree_kind"])
    return dict(pop_target=float(pop_target), epsilon=float(epsilon), node_repeats=int(node_repeats),
                max_attempts=m["max_attempts"], warn_attempts=int(m["warn_attempts"]),
                allow_pair_reselection=m["allow_pair_reselection"], tree_kind=m["tree_kind"],
                cut_policy=policy)


def raise_for_status(code: int, max_attempts, n_pairs=None):
    """Map a per-chain status word to the exception the reference raises at that point."""
    if code == _abi.RC_OK:
        return
    if code == _abi.RC_ERR_MAX_ATTEMPTS:  # tree.py:718
        raise RuntimeError(f"Could not find a possible cut after {max_attempts} attempts.")
    if code == _abi.RC_ERR_ALL_PAIRS_BAD:  # tree_proposals.py:140-144
        raise MetagraphError(
            f"Bipartitioning failed for all {n_pairs} district pairs."
            f"Consider rerunning the chain with a different random seed.")
    if code == _abi.RC_ERR_REVERSIBILITY:  # tree_proposals.py:205-209, :257-261
        raise ReversibilityError("Found more balance edges than the upper bound M allows.")
    if code in (_abi.RC_ERR_NO_CUT_EDGES, _abi.RC_ERR_NO_ROOT):  # random.choice of an empty sequence
        raise IndexError("Cannot choose from an empty sequence")
    from .engine import EngineError

    hint = ""
    if code == _abi.RC_ERR_CAPACITY:
        hint = "merged district pair exceeds the engine's node/edge caps (EngineConfig.cap_nodes / cap_edges)"
    raise EngineError(code, hint)


_WARNING_TEXT = ("\nFailed to find a balanced cut after {} attempts.\n"
                 "If possible, consider enabling pair reselection within your\n"
                 "MarkovChain proposal method to allow the algorithm to select\n"
                 "a different pair of districts for recombination.")


def recom(partition, pop_col: str, pop_target: Union[int, float], epsilon: float, node_repeats: int = 1,
          region_surcharge: Optional[Dict] = None, method: Callable = bipartition_tree):
    """One ReCom proposal (tree_proposals.py:36-146) as one chain-step kernel launch."""
    ctx = partition._ctx
    k = len(partition)
    ekw = engine_kwargs(pop_target, epsilon, node_repeats, region_surcharge, method)
    rkey = tuple(region_surcharge.items()) if region_surcharge else ()
    key = ("recom", rkey) + tuple(sorted((kk, vv) for kk, vv in ekw.items()))
    eng = ctx.engine(key, pop_col, k, region_surcharge, **ekw)  # Bounds are checked by the caller's Validator
    eng, h_assign, h_tally, h_cut = _run_one_step(partition, pop_col, eng, ekw["max_attempts"])
    if not eng._api_warned and int(eng.flags[0].item()) & _abi.RC_FLAG_WARNED:  # tree.py:703-710
        eng._api_warned = True
        warnings.warn(_WARNING_TEXT.format(ekw["warn_attempts"]), BipartitionWarning)
    return partition.__class__._from_device(
        partition, h_assign[0].numpy().copy(), (pop_col,), h_tally[0].numpy().copy(), int(h_cut[0]),
        eng.cut_mask[0].cpu().numpy().view(np.uint32).copy(), eng.last_pair[0].cpu().numpy())


def reversible_engine_kwargs(pop_target, epsilon, M, repeat_until_valid=False) -> Dict:
    """RcConfig fields of a reversible_recom configuration."""
    if repeat_until_valid:
        raise NotImplementedError("reversible_recom(repeat_until_valid=True) has no device path")
    return dict(pop_target=float(pop_target), epsilon=float(epsilon), node_repeats=1, max_attempts=1,
                tree_kind=_abi.RC_TREE_WILSON, cut_policy=_abi.RC_CUT_UNIFORM,
                proposal_kind=_abi.RC_PROPOSAL_REVERSIBLE, rev_M=int(M))


def _run_one_step(partition, pop_col, eng, max_attempts):
    """Host plan in, one device step, host plan out; shared by recom and reversible_recom."""
    k = len(partition)
    bufs = getattr(eng, "_api_bufs", None)
    if bufs is None:
        bufs = eng._api_bufs = eng.host_buffers()
        eng._api_warned = False
    h_assign, h_tally, h_cut, h_status = bufs
    h_assign[0].numpy()[:] = partition._vec
    eng.step_host(h_assign, 1, h_assign, h_tally, h_cut, h_status)
    if int(h_status[0]) == _abi.RC_ERR_CAPACITY:
        # A merged pair larger than the engine's caps (2.4 N / k nodes by default; districts of
        # very uneven node counts exceed it): the reference has no such limit, so the step is
        # taken again on an engine with doubled caps instead of surfacing an engine error.
        bigger = partition._ctx.grow_engine(eng)
        if bigger is not None:
            return _run_one_step(partition, pop_col, bigger, max_attempts)
    raise_for_status(int(h_status[0]), max_attempts, k * (k - 1) / 2)
    return eng, h_assign, h_tally, h_cut


def reversible_recom(partition, pop_col: str, pop_target: Union[int, float], epsilon: float,
                     balance_edge_fn: Callable = tree.find_balanced_edge_cuts_memoization, M: int = 1,
                     repeat_until_valid: bool = False, choice: Callable = None):
    """Reversible ReCom (tree_proposals.py:149-267) as one chain-step kernel launch: uniform
    ordered district pair, one uniform spanning tree, all balanced cuts (more than ``M``:
    ``ReversibilityError``), acceptance with probability ``#cuts / (M * seam length)``.
    Returns ``partition`` itself on a self-loop, like the reference."""
    if balance_edge_fn is not tree.find_balanced_edge_cuts_memoization or choice is not None:
        raise NotImplementedError("reversible_recom hooks (balance_edge_fn, choice) have no device path")
    ctx = partition._ctx
    k = len(partition)
    ekw = reversible_engine_kwargs(pop_target, epsilon, M, repeat_until_valid)
    key = ("reversible",) + tuple(sorted(ekw.items()))
    eng = ctx.engine(key, pop_col, k, None, **ekw)
    eng, h_assign, h_tally, h_cut = _run_one_step(partition, pop_col, eng, 1)
    new_vec = h_assign[0].numpy()
    if np.array_equal(new_vec, partition._vec):
        return partition  # self-loop (no adjacency, no balance edge, or rejected)
    return partition.__class__._from_device(
        partition, new_vec.copy(), (pop_col,), h_tally[0].numpy().copy(), int(h_cut[0]),
        eng.cut_mask[0].cpu().numpy().view(np.uint32).copy(), eng.last_pair[0].cpu().numpy())


class ReCom:
    """tree_proposals.py:270-305 (callable proposal object)."""

    def __init__(self, pop_col: str, ideal_pop: Union[int, float], epsilon: float,
                 method: Callable = bipartition_tree):
        self.pop_col = pop_col
        self.ideal_pop = ideal_pop
        self.epsilon = epsilon
        self.method = method

    def __call__(self, partition):
        return recom(partition, self.pop_col, self.ideal_pop, self.epsilon, method=self.method)
