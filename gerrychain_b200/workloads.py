"""The synthetic workloads BASELINE.json names, as (GraphCSR, seed plan, EngineConfig kwargs).

Seed plans: grids are analytic; the Delaunay plans were produced once by the reference's
``recursive_tree_part`` (oracle/make_seed_plans.py) and are committed under tests/golden/.
"""

from __future__ import annotations

import os
from typing import Tuple

import numpy as np

from . import synthetic as syn
from .config import EngineConfig
from .csr import GraphCSR, assignment_vector, csr_from_edges, csr_from_networkx

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _bounds(csr, k, eps):
    ideal = float(csr.pop.sum()) / k
    # within_percent_of_ideal_population: constraints/validity.py:88-91
    return ideal, (1 - eps) * ideal, (1 + eps) * ideal


def grid_workload(m: int, n: int, k: int, eps: float, n_chains: int, **kw) -> Tuple[GraphCSR, np.ndarray, EngineConfig]:
    g = syn.grid_graph(m, n)
    csr = csr_from_networkx(g, "population")
    plan = syn.grid_quadrants(m, n) if k == 4 else syn.grid_stripes(m, n, k)
    vec, _ = assignment_vector(csr, plan)
    ideal, lo, hi = _bounds(csr, k, eps)
    cfg = EngineConfig(n_districts=k, n_chains=n_chains, pop_target=ideal, epsilon=eps,
                       pop_lower=lo, pop_upper=hi, **kw)
    return csr, vec.astype(np.uint8), cfg


def delaunay_workload(n_nodes: int, k: int, eps: float, n_chains: int, counties: int = 0,
                      surcharge: float = 0.5, pop_range=(500, 2500), **kw):
    # same arrays as csr_from_networkx(syn.delaunay_graph(...)) (tests/test_host_api.py checks
    # it), built from the edge list directly: a 200k-node networkx graph takes half a minute
    rng, pts, e = syn.delaunay_points(n_nodes, seed=0)
    pop = rng.integers(pop_range[0], pop_range[1], n_nodes)
    csr = csr_from_edges(n_nodes, e, pop, node_labels=list(range(n_nodes)))
    regions = None
    if counties:
        from scipy.spatial import cKDTree

        centres = np.random.default_rng(1).random((counties, 2))  # syn.voronoi_counties(seed=1)
        _, idx = cKDTree(centres).query(pts)
        _, first = np.unique(idx, return_index=True)  # codes in order of first appearance
        code = np.empty(counties, dtype=np.int32)
        code[idx[np.sort(first)]] = np.arange(len(first), dtype=np.int32)
        regions = {"county": surcharge}
        csr = csr.with_regions(code[idx][None, :], [surcharge], ("county",))
    path = os.path.join(_GOLDEN, f"plan_delaunay{n_nodes}_k{k}.npy")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: run oracle/make_seed_plans.py in the build container")
    vec = np.load(path)
    ideal, lo, hi = _bounds(csr, k, eps)
    from .config import cut_policy_for
    from . import _abi

    cfg = EngineConfig(n_districts=k, n_chains=n_chains, pop_target=ideal, epsilon=eps,
                       pop_lower=lo, pop_upper=hi,
                       cut_policy=cut_policy_for(regions, kw.get("tree_kind", _abi.RC_TREE_MST)), **kw)
    return csr, vec.astype(np.uint8), cfg


def many_districts(m: int = 60, n: int = 60, bh: int = 3, bw: int = 4, eps: float = 0.1, n_chains: int = 1, **kw):
    """An m x n grid cut into bh x bw blocks: (m / bh) * (n / bw) small districts (300 by
    default) - the shape of a plan with more than 255 districts (16-bit district labels)."""
    if m % bh or n % bw:
        raise ValueError("the blocks must tile the grid")
    g = syn.grid_graph(m, n)
    csr = csr_from_networkx(g, "population")
    plan = {(i, j): (i // bh) * (n // bw) + j // bw for i in range(m) for j in range(n)}
    vec, _ = assignment_vector(csr, plan)
    k = (m // bh) * (n // bw)
    ideal, lo, hi = _bounds(csr, k, eps)
    cfg = EngineConfig(n_districts=k, n_chains=n_chains, pop_target=ideal, epsilon=eps,
                       pop_lower=lo, pop_upper=hi, **kw)
    return csr, vec.astype(np.uint8 if k <= 255 else np.uint16), cfg


def config1(n_chains=1, **kw):
    """Grid 20x20, 4 districts, eps 0.05 (BASELINE.json configs[0])."""
    return grid_workload(20, 20, 4, 0.05, n_chains, **kw)


def config2(n_chains=1024, **kw):
    """Grid 100x100, 10 districts, eps 0.05, MST (configs[1])."""
    return grid_workload(100, 100, 10, 0.05, n_chains, **kw)


def config3(n_chains=1024, **kw):
    """Delaunay 9k nodes, 18 districts, eps 0.02 (configs[2]; n_chains is per GPU)."""
    return delaunay_workload(9000, 18, 0.02, n_chains, **kw)


def config5(n_chains=1024, **kw):
    """configs[4]: config 3 + 67 Voronoi counties, region_surcharge={"county": 0.5}."""
    return delaunay_workload(9000, 18, 0.02, n_chains, counties=67, **kw)


def config4(n_chains=1024, **kw):
    """configs[3]: Delaunay 200k nodes, 50 districts, eps 0.05, uniform_spanning_tree (Wilson).
    Merged pairs (~8k nodes) exceed an SM's shared memory: the engine runs in spill mode."""
    from . import _abi

    kw.setdefault("tree_kind", _abi.RC_TREE_WILSON)
    return delaunay_workload(200000, 50, 0.05, n_chains, pop_range=(20, 200), **kw)
