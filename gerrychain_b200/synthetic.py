"""Synthetic dual graphs of the shapes BASELINE.json names.

* Grid m x n: same node set / attributes as the reference's ``create_grid_graph``
  (/root/reference/gerrychain/grid.py:142-181): nodes ``(i, j)``, ``population = 1``.
* Planar Delaunay triangulation of uniform random points with integer populations.
* Voronoi "counties" for the region-aware configuration.
"""

from __future__ import annotations

from typing import Dict, Hashable, Tuple

import networkx
import numpy as np


def grid_graph(m: int, n: int, with_diagonals: bool = False) -> networkx.Graph:
    g = networkx.grid_2d_graph(m, n)
    networkx.set_edge_attributes(g, 1, "shared_perim")
    if with_diagonals:
        diag = [((i, j), (i + 1, j + 1)) for i in range(m - 1) for j in range(n - 1)]
        diag += [((i, j + 1), (i + 1, j)) for i in range(m - 1) for j in range(n - 1)]
        g.add_edges_from(diag, shared_perim=0)
    networkx.set_node_attributes(g, 1, "population")
    networkx.set_node_attributes(g, 1, "area")
    for (i, j) in g.nodes:  # grid.py:196-236: sides of the cell on the rim; corners have two
        sides = int(i in (0, m - 1)) + int(j in (0, n - 1))
        g.nodes[(i, j)]["boundary_node"] = sides > 0
        if sides:
            g.nodes[(i, j)]["boundary_perim"] = sides
    return g


def grid_quadrants(m: int, n: int) -> Dict[Hashable, int]:
    """Default ``Grid`` seed plan (/root/reference/gerrychain/grid.py:267-288)."""
    tx, ty = m // 2, n // 2
    return {
        (i, j): (0 if i < tx else 1) + (0 if j < ty else 2)
        for i in range(m)
        for j in range(n)
    }


def grid_stripes(m: int, n: int, k: int) -> Dict[Hashable, int]:
    """k equal row-stripes: ``district = i // (m / k)``."""
    if m % k:
        raise ValueError("m must be divisible by k")
    w = m // k
    return {(i, j): i // w for i in range(m) for j in range(n)}


def delaunay_points(n_nodes: int, seed: int = 0):
    """Points, unique Delaunay edges and integer populations, all from one rng."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    pts = rng.random((n_nodes, 2))
    tri = Delaunay(pts)
    s = tri.simplices
    e = np.concatenate([s[:, [0, 1]], s[:, [1, 2]], s[:, [0, 2]]])
    e = np.unique(np.sort(e, axis=1), axis=0)
    return rng, pts, e


def delaunay_graph(
    n_nodes: int, seed: int = 0, pop_range: Tuple[int, int] = (500, 2500)
) -> networkx.Graph:
    rng, pts, e = delaunay_points(n_nodes, seed)
    pop = rng.integers(pop_range[0], pop_range[1], n_nodes)
    g = networkx.Graph()
    for i in range(n_nodes):
        g.add_node(i, population=int(pop[i]), x=float(pts[i, 0]), y=float(pts[i, 1]))
    g.add_edges_from((int(a), int(b)) for a, b in e)
    return g


def voronoi_counties(g: networkx.Graph, n_counties: int, seed: int = 1, key="county"):
    """Tag every node with the nearest of ``n_counties`` random centres."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    centres = rng.random((n_counties, 2))
    nodes = list(g.nodes)
    xy = np.array([(g.nodes[v]["x"], g.nodes[v]["y"]) for v in nodes])
    _, idx = cKDTree(centres).query(xy)
    for v, c in zip(nodes, idx):
        g.nodes[v][key] = int(c)
    return g


def stripes_by_x(g: networkx.Graph, k: int, pop_col: str = "population"):
    """Crude connected-ish seed plan: cut the x-sorted node list into k equal-pop runs.

    Only used for smoke tests; real seed plans come from ``recursive_tree_part``.
    """
    nodes = sorted(g.nodes, key=lambda v: g.nodes[v]["x"])
    pops = np.array([g.nodes[v][pop_col] for v in nodes], dtype=np.float64)
    cum = np.cumsum(pops)
    lab = np.minimum((cum - 1e-9) * k // cum[-1], k - 1).astype(int)
    return {v: int(d) for v, d in zip(nodes, lab)}
