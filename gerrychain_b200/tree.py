"""Names of the reference's tree module that select engine behaviour.

In the reference these are the algorithms themselves (/root/reference/gerrychain/tree.py);
here they are *selectors*: ``recom(method=partial(bipartition_tree, spanning_tree_fn=
uniform_spanning_tree, max_attempts=..., allow_pair_reselection=True))`` is read by the
proposal and turned into the engine configuration (RcConfig).  Calling one of them
directly raises: the algorithms only exist as CUDA kernels (no CPU fallback).
"""

from __future__ import annotations


class BipartitionWarning(UserWarning):
    """tree.py:427-432."""


class ReselectException(Exception):
    """tree.py:435-443."""


class BalanceError(Exception):
    """tree.py:1441-1442."""


class PopulationBalanceError(Exception):
    """tree.py:1445-1446."""


def _selector_only(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"gerrychain_b200.tree.{name} is a selector for the CUDA engine (pass it to recom / "
            "bipartition_tree as in GerryChain); it has no host implementation")
    fn.__name__ = name
    fn.__qualname__ = name
    return fn


random_spanning_tree = _selector_only("random_spanning_tree")  # tree.py:60-94  -> RC_TREE_MST
uniform_spanning_tree = _selector_only("uniform_spanning_tree")  # tree.py:97-132 -> RC_TREE_WILSON
find_balanced_edge_cuts_memoization = _selector_only("find_balanced_edge_cuts_memoization")  # tree.py:351
# tree.py:243-286: the leaf-contraction variant finds the SAME set of balanced cuts (both test
# has_ideal_population on every non-root node of the rooted tree) and the cut is then drawn
# uniformly from it, so it selects the same device path (the order of the list differs, which
# only matters to a replay of recorded choices)
find_balanced_edge_cuts_contraction = _selector_only("find_balanced_edge_cuts_contraction")
_region_preferred_max_weight_choice = _selector_only("_region_preferred_max_weight_choice")  # tree.py:501
_max_weight_choice = _selector_only("_max_weight_choice")  # tree.py:446


def bipartition_tree(graph=None, pop_col=None, pop_target=None, epsilon=None, node_repeats=1,
                     spanning_tree=None, spanning_tree_fn=random_spanning_tree, region_surcharge=None,
                     balance_edge_fn=find_balanced_edge_cuts_memoization, one_sided_cut=False,
                     choice=None, max_attempts=100000, warn_attempts=1000,
                     allow_pair_reselection=False, cut_choice=_region_preferred_max_weight_choice):
    """Signature of tree.py:581-597; used through ``functools.partial`` as ``recom``'s
    ``method``.  See module docstring."""
    raise NotImplementedError(
        "bipartition_tree runs inside the CUDA chain-step kernel; use it as recom(method=...)")


#: keyword arguments of bipartition_tree the engine implements, and their defaults
ENGINE_KWARGS = dict(spanning_tree_fn=random_spanning_tree, max_attempts=100000, warn_attempts=1000,
                     allow_pair_reselection=False, cut_choice=_region_preferred_max_weight_choice,
                     balance_edge_fn=find_balanced_edge_cuts_memoization, one_sided_cut=False)


def recursive_tree_part(graph, parts, pop_target, pop_col, epsilon, node_repeats=1, method=None):
    """tree.py:971-1085 on the device (see gerrychain_b200.seed)."""
    from .seed import recursive_tree_part as _impl

    return _impl(graph, parts, pop_target, pop_col, epsilon, node_repeats=node_repeats, method=method)
