"""Loads the recorded reference steps (tests/golden/replay_*.npz, written by
oracle/record_reference.py) into the flat arrays the replay ABI takes."""

from __future__ import annotations

import glob
import json
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from gerrychain_b200 import _abi
from gerrychain_b200.config import EngineConfig, cut_policy_for
from gerrychain_b200.csr import csr_from_edges

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STEP_DTYPE = np.dtype(
    [("pair_edge", "<i4"), ("n_trees", "<i4"), ("first_tree", "<i8"), ("root", "<i4"),
     ("cut_edge", "<i4")], align=True)
assert STEP_DTYPE.itemsize == 24


@dataclass
class ReplayArrays:
    steps: np.ndarray
    weights: Optional[np.ndarray] = None  # [T][E] f64
    weight_rank: Optional[np.ndarray] = None  # [T][E] u32, order of weights (ties: edge id)
    wilson_root: Optional[np.ndarray] = None  # [T]
    stack_ptr: Optional[np.ndarray] = None  # [T][N+1] i64
    stack_val: Optional[np.ndarray] = None


@dataclass
class Golden:
    name: str
    meta: dict
    csr: object
    cfg: EngineConfig
    init_assignment: np.ndarray
    rec: ReplayArrays
    assignment: np.ndarray  # [S][N] reference assignment after each step
    n_cut_edges: np.ndarray
    population: np.ndarray
    balanced_ptr: np.ndarray
    balanced: np.ndarray
    tree_ptr: np.ndarray


def names(kind=None):
    out = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "replay_*.npz")))
    if kind == "mst":
        out = [n for n in out if "wilson" not in n]
    elif kind == "wilson":
        out = [n for n in out if "wilson" in n]
    return out


def load(name, inject_cut=True) -> Golden:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    if "graph" in meta:
        # big fixtures name the generator call instead of storing the graph
        # (oracle/record_config_size.py); the checksum pins what the recorder saw
        import zlib

        from gerrychain_b200 import synthetic

        kind, n, gseed, lo, hi = meta["graph"]
        assert kind == "delaunay"
        rng, _, e = synthetic.delaunay_points(n, gseed)
        pop = rng.integers(lo, hi, n).astype(np.int32)
        csr = csr_from_edges(n, e.astype(np.int32), pop)
        crc = zlib.crc32(np.ascontiguousarray(csr.pop, dtype=np.int32).tobytes(),
                         zlib.crc32(np.ascontiguousarray(csr.edge_uv, dtype=np.int32).tobytes()))
        assert crc == meta["graph_crc32"], "regenerated graph differs from the recorded one"
        n_nodes = n
    else:
        n_nodes = int(z["pop"].shape[0])
        csr = csr_from_edges(n_nodes, z["edge_uv"], z["pop"])
        assert np.array_equal(csr.edge_uv, z["edge_uv"])
    regions = None
    if z["region_ids"].size:
        csr = csr.with_regions(z["region_ids"], z["surcharges"], meta["region_names"])
        regions = dict(zip(meta["region_names"], z["surcharges"].tolist()))
    tree_kind = _abi.RC_TREE_MST if meta["tree"] == "mst" else _abi.RC_TREE_WILSON
    cfg = EngineConfig(
        n_districts=meta["k"], n_chains=1, pop_target=meta["pop_target"],
        epsilon=meta["epsilon"], node_repeats=meta["node_repeats"], tree_kind=tree_kind,
        cut_policy=cut_policy_for(regions, tree_kind), seed=0)
    S = meta["n_steps"]
    tree_ptr = z["tree_ptr"]
    steps = np.zeros(S, dtype=STEP_DTYPE)
    steps["pair_edge"] = z["pair_edge"]
    steps["n_trees"] = np.diff(tree_ptr)
    steps["first_tree"] = tree_ptr[:-1]
    steps["root"] = z["root"]
    steps["cut_edge"] = z["cut_edge"] if inject_cut else -1
    T = int(tree_ptr[-1])
    rec = ReplayArrays(steps=steps)
    if meta["tree"] == "mst":
        w = np.full((T, csr.n_edges), np.nan, dtype=np.float64)
        wp = z["w_ptr"]
        t_of = np.repeat(np.arange(T), np.diff(wp))
        w[t_of, z["w_eid"]] = z["w_val"]
        rec.weights = w
        # NaN (edge outside the merged pair) sorts last; stable -> ties fall to the lower id
        order = np.argsort(w, axis=1, kind="stable")
        rank = np.empty_like(order)
        np.put_along_axis(rank, order, np.broadcast_to(np.arange(csr.n_edges), order.shape), axis=1)
        rec.weight_rank = rank.astype(np.uint32)
    else:
        rec.wilson_root = z["wilson_root"].astype(np.int32)
        sp = np.zeros((T, n_nodes + 1), dtype=np.int64)
        wn_ptr, wn_node, ws_ptr = z["wn_ptr"], z["wn_node"], z["ws_ptr"]
        for t in range(T):
            lo, hi = int(wn_ptr[t]), int(wn_ptr[t + 1])
            cnt = np.zeros(n_nodes, dtype=np.int64)
            cnt[wn_node[lo:hi]] = ws_ptr[lo + 1:hi + 1] - ws_ptr[lo:hi]
            # stacks of one tree are stored contiguously in ascending node order
            sp[t, 0] = ws_ptr[lo]
            sp[t, 1:] = ws_ptr[lo] + np.cumsum(cnt)
        rec.stack_ptr = sp
        rec.stack_val = z["ws_val"].astype(np.int32)
    return Golden(name, meta, csr, cfg, z["init_assignment"], rec, z["assignment"],
                  z["n_cut_edges"], z["population"], z["balanced_ptr"], z["balanced"], tree_ptr)


def seedplan_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "seedplan_*.npz")))


def load_seedplan(name):
    """Recorded reference recursive_tree_part (oracle/record_seed_plans.py):
    (csr, cfg, ReplayArrays, reference plan, cuts per split)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    csr = csr_from_edges(int(z["pop"].shape[0]), z["edge_uv"], z["pop"])
    cfg = EngineConfig(n_districts=meta["k"], n_chains=1, pop_target=meta["pop_target"],
                       epsilon=meta["epsilon"], seed=0)
    tree_ptr = z["tree_ptr"]
    S = len(tree_ptr) - 1
    steps = np.zeros(S, dtype=STEP_DTYPE)
    steps["pair_edge"] = -1
    steps["n_trees"] = np.diff(tree_ptr)
    steps["first_tree"] = tree_ptr[:-1]
    steps["root"] = z["root"]
    steps["cut_edge"] = z["cut_edge"]
    T = int(tree_ptr[-1])
    w = np.full((T, csr.n_edges), np.nan, dtype=np.float64)
    wp = z["w_ptr"]
    t_of = np.repeat(np.arange(T), np.diff(wp))
    w[t_of, z["w_eid"]] = z["w_val"]
    order = np.argsort(w, axis=1, kind="stable")
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.broadcast_to(np.arange(csr.n_edges), order.shape), axis=1)
    rec = ReplayArrays(steps=steps, weights=w, weight_rank=rank.astype(np.uint32))
    return csr, cfg, rec, z["plan"], z["n_cuts"]
