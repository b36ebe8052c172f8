"""Compactness updaters of ``GeographicPartition`` / ``Grid`` (perimeter, exterior / interior
boundaries, boundary nodes, Polsby-Popper) against what the unmodified reference reports for
every state of one of its own ReCom chains (tests/golden/compactness_*.npz, recorded by
oracle/record_compactness.py; the reference's values went through its incremental flow
updaters, updaters/compactness.py).  Integer geometry (the reference's Grid) must match
exactly; float geometry within 1e-9 relative + 1e-12 absolute (the reference adds and removes
perimeters step by step - a district that lost its last boundary node keeps a 2e-16 residue
there - while here each state is summed afresh)."""

import json
import os

import networkx as nx
import numpy as np
import pytest

import _golden
from gerrychain_b200 import GeographicPartition, Grid, metrics, updaters

CASES = ["grid12", "delaunay300"]


def _load(name):
    z = np.load(os.path.join(_golden.GOLDEN, f"compactness_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    labels = [tuple(v) if isinstance(v, list) else v for v in meta["labels"]]
    integral = bool(z["integral"])
    num = (lambda x: int(x)) if integral else float
    g = nx.Graph()
    for i, lab in enumerate(labels):
        g.add_node(lab, population=int(z["node_population"][i]), area=num(z["node_area"][i]),
                   boundary_node=bool(z["node_boundary"][i]))
        if z["node_boundary"][i]:
            g.nodes[lab]["boundary_perim"] = num(z["node_boundary_perim"][i])
    for (u, v), w in zip(z["edge_uv"].tolist(), z["edge_shared_perim"].tolist()):
        g.add_edge(labels[u], labels[v], shared_perim=num(w))
    return z, meta, labels, g, integral


@pytest.mark.parametrize("name", CASES)
def test_compactness_updaters_match_reference(name):
    z, meta, labels, g, integral = _load(name)
    parts = meta["parts"]
    for t in range(z["plan"].shape[0]):
        plan = {lab: parts[int(d)] for lab, d in zip(labels, z["plan"][t])}
        p = GeographicPartition(g, plan)
        ext, inte, per = p["exterior_boundaries"], p["interior_boundaries"], p["perimeter"]
        got = np.array([[ext[q], inte[q], per[q]] for q in parts], dtype=np.float64)
        want = np.stack([z["exterior"][t], z["interior"][t], z["perimeter"][t]], axis=1)
        if integral:
            assert all(isinstance(ext[q], int) and isinstance(inte[q], int) for q in parts)
            assert np.array_equal(got, want), t
        else:
            assert np.allclose(got, want, rtol=1e-9, atol=1e-12), t
        pp = np.array([metrics.compute_polsby_popper(z["area"][t][i], per[q]) for i, q in enumerate(parts)])
        assert np.allclose(pp, z["polsby"][t], rtol=1e-9, atol=0.0)
    assert p["boundary_nodes"] == {labels[i] for i in z["boundary_nodes"].tolist()}
    sets = updaters.exterior_boundaries_as_a_set(p)
    assert set().union(*sets.values()) == p["boundary_nodes"]
    assert all(plan[v] == q for q, members in sets.items() for v in members)
    assert updaters.perimeter_of_part(p, parts[0]) == per[parts[0]]


def test_default_updaters_are_the_reference_sets():
    # partition/geographic.py:21-29 and grid.py:44-54
    assert set(GeographicPartition.default_updaters) == {
        "perimeter", "exterior_boundaries", "interior_boundaries", "boundary_nodes", "cut_edges", "area",
        "cut_edges_by_part"}
    assert set(Grid.default_updaters) == {
        "cut_edges", "population", "perimeter", "exterior_boundaries", "interior_boundaries", "boundary_nodes",
        "area", "polsby_popper", "cut_edges_by_part"}


def test_grid_graph_carries_the_reference_geometry():
    """create_grid_graph / tag_boundary_nodes (grid.py:142-236): the fixture's graph is the
    reference's own Grid((12, 12))."""
    z, meta, labels, g, _ = _load("grid12")
    ours = Grid((12, 12)).graph
    assert list(ours.nodes) == labels
    for lab in labels:
        assert dict(ours.nodes[lab]) == dict(g.nodes[lab]), lab
    assert {tuple(sorted(e)) for e in ours.edges} == {tuple(sorted(e)) for e in g.edges}
    assert all(ours.edges[e]["shared_perim"] == g.edges[e]["shared_perim"] for e in g.edges)


def test_polsby_popper_of_a_district_without_perimeter_is_nan():
    assert np.isnan(metrics.compute_polsby_popper(1.0, 0))  # metrics/compactness.py:17-20


@pytest.mark.gpu
def test_grid_polsby_popper_and_area_on_the_device_path():
    """``area`` is a device Tally; the whole default updater set of a Grid state."""
    z, meta, labels, g, _ = _load("grid12")
    parts = meta["parts"]
    t = z["plan"].shape[0] - 1
    grid = Grid((12, 12), assignment={lab: parts[int(d)] for lab, d in zip(labels, z["plan"][t])})
    assert [grid["area"][q] for q in parts] == z["area"][t].tolist()
    assert np.allclose([grid["polsby_popper"][q] for q in parts], z["polsby"][t], rtol=1e-12)
    by_part = grid["cut_edges_by_part"]
    assert [sum(grid.graph.edges[e]["shared_perim"] for e in by_part[q]) for q in parts] == z["interior"][t].tolist()
