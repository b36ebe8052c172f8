"""gerrychain_b200 - a B200-native batched ReCom engine behind GerryChain's API.

Public surface (same names and call signatures as mggg/GerryChain for the ReCom path):
``MarkovChain``, ``Partition`` / ``GeographicPartition``, ``proposals.recom`` / ``ReCom``,
``constraints.within_percent_of_ideal_population`` / ``Validator`` / ``Bounds``,
``updaters.Tally`` / ``cut_edges``, ``accept.always_accept``, ``Grid`` - plus
``BatchedMarkovChain``, the multi-chain runner the reference does not have.

Every compute step runs in the in-tree CUDA library (``librecom_b200.so``, sm_100a) through
a ctypes C ABI (include/recom_b200.h).  There is no CPU fallback: using the package without
the built library or without a CUDA device raises.
"""

from . import accept, constraints, metrics, proposals, tree, updaters  # noqa: F401
from .accept import always_accept
from .chain import BatchedMarkovChain, ChainBatch, MarkovChain
from .constraints import Bounds, Validator, within_percent_of_ideal_population
from .graph import Graph
from .grid import Grid
from .partition import Assignment, GeographicPartition, Partition
from .record import Record, Replay
from .proposals import MetagraphError, ReCom, ReversibilityError, recom, reversible_recom
from .tree import BipartitionWarning, bipartition_tree
from .updaters import Election, Tally, cut_edges, tally_region_splits

__all__ = [
    "MarkovChain", "BatchedMarkovChain", "ChainBatch", "Partition", "GeographicPartition", "Assignment",
    "Grid", "Graph", "recom", "reversible_recom", "ReversibilityError", "ReCom", "MetagraphError", "BipartitionWarning", "bipartition_tree",
    "within_percent_of_ideal_population", "Validator", "Bounds", "Tally", "cut_edges",
    "Election", "tally_region_splits", "Record", "Replay", "always_accept", "accept", "constraints", "metrics", "proposals", "tree", "updaters",
]
