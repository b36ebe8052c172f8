"""GPU tests of the GerryChain-shaped API: the reference's own hot-path tests, run against
the CUDA engine: recom as a proposal keeps districts contiguous
(reference tests/test_tree.py:291-302), incremental cut_edges / Tally == naive recomputation
at every step (tests/updaters/test_cut_edges.py:128-161, tests/test_tally.py:56-76), the
chain's step accounting (tests/test_chain.py), the max_attempts error and warn_attempts
warning (tests/test_region_aware.py:102-112, :236-250), and the batched runner against the
oracle."""

import functools
import random
import warnings

import networkx
import numpy as np
import pytest

import gerrychain_b200 as gc
from gerrychain_b200 import Grid, MarkovChain, Partition, Tally, cut_edges, recom
from gerrychain_b200.accept import always_accept
from gerrychain_b200.constraints import within_percent_of_ideal_population

pytestmark = pytest.mark.gpu


def _contiguous(partition):
    g = partition.graph
    return all(networkx.is_connected(g.subgraph(nodes)) for nodes in partition.parts.values())


def _naive_cut_edges(partition):
    m = partition.assignment.mapping
    return {tuple(sorted(e)) for e in partition.graph.edges if m[e[0]] != m[e[1]]}


def _naive_tally(partition, field):
    out = {}
    for node, part in partition.assignment.items():
        out[part] = out.get(part, 0) + partition.graph.nodes[node][field]
    return out


def _three_by_three():
    g = networkx.grid_2d_graph(3, 3)
    g = networkx.convert_node_labels_to_integers(g)
    for n in g.nodes:
        g.nodes[n]["pop"] = 1
    return g


def test_recom_works_as_a_proposal():
    graph = _three_by_three()
    partition = Partition(graph, {0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 2, 6: 2, 7: 2, 8: 2}, {"pop": Tally("pop")})
    ideal = sum(partition["pop"].values()) / 2
    proposal = functools.partial(recom, pop_col="pop", pop_target=ideal, epsilon=0.25, node_repeats=5)
    constraints = [within_percent_of_ideal_population(partition, 0.25, "pop")]
    chain = MarkovChain(proposal, constraints, always_accept, partition, total_steps=100)
    n = 0
    for state in chain:
        assert _contiguous(state)
        assert set(state["pop"].values()) <= {4, 5}
        n += 1
    assert n == 100


def test_grid_chain_incremental_updaters_equal_naive():
    random.seed(2024)
    grid = Grid((10, 10), with_diagonals=True)
    ideal = sum(grid["population"].values()) / len(grid)
    assert grid["cut_edges"] == _naive_cut_edges(grid)
    proposal = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.5, node_repeats=1)
    chain = MarkovChain(proposal, [within_percent_of_ideal_population(grid, 0.5)], always_accept, grid, 60)
    states = []
    for state in chain:
        assert state["cut_edges"] == _naive_cut_edges(state)
        assert state["population"] == _naive_tally(state, "population")
        assert state["area"] == _naive_tally(state, "area")
        assert len(state["cut_edges"]) == state.n_cut_edges
        assert _contiguous(state)
        states.append(state)
    assert states[0] is grid and len(states) == 60
    assert isinstance(states[-1], Grid) and states[-1].dimensions == (10, 10)
    child = states[-1]
    assert child.parent is not None and child.flips
    merged = set(child.flips)
    touched = {child.parent.assignment.mapping[n] for n in merged}
    assert len(touched) == 2  # flips list every node of the two merged districts (tree.py:947-961)
    assert merged == set().union(*(child.parent.parts[d] for d in touched))
    assert len({tuple(s.assignment_vector()) for s in states}) > 10


def test_flip_and_first_time_scans():
    grid = Grid((6, 6))
    flipped = grid.flip({(2, 2): 1, (2, 3): 1})
    assert flipped["cut_edges"] == _naive_cut_edges(flipped)
    assert flipped["population"] == _naive_tally(flipped, "population")
    assert flipped.crosses_parts(((2, 2), (2, 1))) != grid.crosses_parts(((2, 2), (2, 1)))
    assert repr(grid).startswith("6x6 Grid")
    with pytest.raises(KeyError):
        Partition(grid.graph, {(0, 0): 1}, {})
    with pytest.raises(TypeError):
        Partition(42, {}, {})


def test_same_seed_same_chain():
    def run(seed):
        random.seed(seed)
        grid = Grid((12, 12))
        ideal = sum(grid["population"].values()) / 4
        p = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.1)
        chain = MarkovChain(p, [within_percent_of_ideal_population(grid, 0.1)], always_accept, grid, 30)
        return [s.assignment_vector().tobytes() for s in chain]

    assert run(7) == run(7)
    assert run(7) != run(8)


def test_max_attempts_raises_like_the_reference():
    random.seed(0)
    grid = Grid((20, 20))
    ideal = sum(grid["population"].values()) / 4
    method = functools.partial(gc.bipartition_tree, max_attempts=1)
    p = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.01, method=method)
    chain = MarkovChain(p, [within_percent_of_ideal_population(grid, 0.01)], always_accept, grid, 200)
    with pytest.raises(RuntimeError, match="Could not find a possible cut after 1 attempts."):
        for _ in chain:
            pass


def test_warn_attempts_warns_like_the_reference():
    random.seed(0)
    grid = Grid((20, 20))
    ideal = sum(grid["population"].values()) / 4
    method = functools.partial(gc.bipartition_tree, warn_attempts=2)
    p = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.01, method=method)
    chain = MarkovChain(p, [within_percent_of_ideal_population(grid, 0.01)], always_accept, grid, 60)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for _ in chain:
            pass
    assert any(issubclass(x.category, gc.BipartitionWarning) and
               "Failed to find a balanced cut after 2 attempts." in str(x.message) for x in w)


def test_region_aware_recom_through_the_api():
    random.seed(1)
    g = networkx.grid_2d_graph(8, 8)
    for (i, j) in g.nodes:
        g.nodes[(i, j)].update(TOTPOP=1, county=i // 4 * 2 + j // 4, district=i // 2)
    part = Partition(g, "district", {"population": Tally("TOTPOP", alias="population"), "cut_edges": cut_edges})
    p = functools.partial(recom, pop_col="TOTPOP", pop_target=16, epsilon=0.1, region_surcharge={"county": 2.0})
    chain = MarkovChain(p, [within_percent_of_ideal_population(part, 0.1)], always_accept, part, 80)
    last = None
    for state in chain:
        assert _contiguous(state)
        last = state
    # a surcharge of 2.0 keeps trees inside counties: few county-splitting edges are cut
    county = {n: g.nodes[n]["county"] for n in g.nodes}
    splits = sum(len({county[n] for n in nodes}) > 1 for nodes in last.parts.values())
    assert splits <= 2


def test_batched_markov_chain_matches_oracle_and_api():
    from oracle import oracle
    from gerrychain_b200 import BatchedMarkovChain
    from gerrychain_b200.config import EngineConfig

    grid = Grid((20, 20))
    ideal = sum(grid["population"].values()) / 4
    p = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.05)
    cons = [within_percent_of_ideal_population(grid, 0.05)]
    runner = BatchedMarkovChain(p, cons, always_accept, grid, total_steps=41, n_chains=48, seed=99)
    batches = list(runner)
    assert len(batches) == 41 and batches[0].step == 1 and batches[-1].step == 41
    last = batches[-1]
    eng = runner.engine
    cfg = EngineConfig(n_districts=4, n_chains=48, pop_target=ideal, epsilon=0.05, seed=99,
                       pop_lower=cons[0].bounds[0], pop_upper=cons[0].bounds[1])
    ora = oracle.OracleChains(eng.csr, cfg, grid.assignment_vector()).run(40, n_threads=4)
    assert np.array_equal(last.assignment.cpu().numpy(), ora.assignments())
    assert np.array_equal(last.cut_edge_counts(), ora.cut_count)
    assert np.array_equal(last.population(), ora.tally)
    part = last.partition(5)
    assert part["cut_edges"] == _naive_cut_edges(part)
    assert part["population"] == _naive_tally(part, "population")
    assert len(part["cut_edges"]) == int(ora.cut_count[5])
    votes = {n: (i * 7) % 13 for i, n in enumerate(grid.graph.nodes)}
    networkx.set_node_attributes(grid.graph, votes, "votes")
    got = last.tally("votes")
    a = last.assignment.cpu().numpy()
    col = np.array([votes[n] for n in grid.graph.nodes])
    want = np.stack([np.bincount(a[c], weights=col, minlength=4) for c in range(48)]).astype(np.int64)
    assert np.array_equal(got, want)
    cut, dev = last.histograms(256, 64, 0.001)
    assert int(cut.sum()) == 48 and int(dev.sum()) == 48
    # run_to_end from a fresh iterator gives the same ensemble in one launch
    runner2 = BatchedMarkovChain(p, cons, always_accept, grid, total_steps=41, n_chains=48, seed=99)
    assert np.array_equal(runner2.run_to_end().assignment.cpu().numpy(), ora.assignments())
    with pytest.raises(TypeError):
        BatchedMarkovChain(p, cons, lambda s: True, grid, 10, 4)


def test_from_random_assignment_and_recursive_tree_part():
    random.seed(11)
    g = networkx.grid_2d_graph(12, 12)
    for n in g.nodes:
        g.nodes[n]["population"] = 1
    part = Partition.from_random_assignment(g, 6, 0.1, "population", updaters={"population": Tally("population")})
    assert len(part) == 6 and _contiguous(part)
    assert all(abs(p - 24) <= 2.4 for p in part["population"].values())
    plan = gc.tree.recursive_tree_part(g, ["a", "b", "c"], 48, "population", 0.1)
    assert set(plan.values()) == {"a", "b", "c"} and set(plan) == set(g.nodes)
    from gerrychain_b200.seed import random_plans

    plans = random_plans(g, 4, 0.1, "population", n_plans=32, seed=3)
    assert plans.shape == (32, 144) and len({p.tobytes() for p in plans}) > 16
    # seed plans feed the batched runner
    ideal = 36
    p = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.1)
    init = Partition(g, {n: int(plans[0][i]) for i, n in enumerate(g.nodes)}, {"population": Tally("population")})
    runner = gc.BatchedMarkovChain(p, [within_percent_of_ideal_population(init, 0.1)], always_accept, init,
                                   total_steps=11, n_chains=32, seed=1, initial_assignments=plans)
    last = runner.run_to_end()
    assert (np.abs(last.population() - ideal) <= 3.6).all()


def test_reversible_recom_as_a_proposal():
    """Reference tests/test_tree.py:305-317: a reversible_recom chain keeps districts contiguous."""
    random.seed(5)
    graph = _three_by_three()
    partition = Partition(graph, {0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 2, 6: 2, 7: 2, 8: 2}, {"pop": Tally("pop")})
    ideal = sum(partition["pop"].values()) / 2
    proposal = functools.partial(gc.reversible_recom, pop_col="pop", pop_target=ideal, epsilon=0.25, M=4)
    chain = MarkovChain(proposal, [within_percent_of_ideal_population(partition, 0.25, "pop")], always_accept,
                        partition, total_steps=300)
    seen = set()
    for state in chain:
        assert _contiguous(state)
        seen.add(state.assignment_vector().tobytes())
    assert len(seen) > 1
    runner = gc.BatchedMarkovChain(proposal, [within_percent_of_ideal_population(partition, 0.25, "pop")],
                                   always_accept, partition, total_steps=201, n_chains=16, seed=2)
    assert (np.abs(runner.run_to_end().population() - ideal) <= 0.25 * ideal).all()


def test_flip_chain_incremental_updaters_equal_naive():
    """Reference tests/test_tally.py:56-76 and tests/updaters/test_cut_edges.py:31-125: after
    every (multi-node) flip the device-backed Tally / cut_edges equal the naive recomputation,
    and no edge appears as both (u, v) and (v, u)."""
    rng = random.Random(4)
    grid = Grid((8, 8))
    state = grid
    nodes = list(grid.graph.nodes)
    for _ in range(60):
        flips = {}
        for _ in range(rng.randint(1, 3)):
            e = rng.choice(sorted(state["cut_edges"]))
            src, dst = (e[0], e[1]) if rng.random() < 0.5 else (e[1], e[0])
            flips[src] = state.assignment.mapping[dst]
        state = state.flip(flips)
        assert state["cut_edges"] == _naive_cut_edges(state)
        assert state["population"] == _naive_tally(state, "population")
        assert not any((v, u) in state["cut_edges"] for (u, v) in state["cut_edges"])
        assert state.flips == flips and state.parent is not None
    assert len(nodes) == 64


def _stats_fixture_graph(name):
    import json
    import os

    import _golden

    z = np.load(os.path.join(_golden.GOLDEN, f"stats_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    g = networkx.Graph()
    for v in range(z["pop"].shape[0]):
        attrs = {"pop": int(z["pop"][v]), "votes_a": int(z["votes"][0, v]), "votes_b": int(z["votes"][1, v])}
        for c, col in enumerate(meta["region_cols"]):
            attrs[col] = int(z["region_ids"][c, v])
        g.add_node(v, **attrs)
    g.add_edges_from((int(a), int(b)) for a, b in z["edge_uv"])
    return z, meta, g


@pytest.mark.parametrize("name", ["8x8_muni", "delaunay600"])
def test_statistics_updaters_match_reference_fixture(name):
    """tally_region_splits (reference tests/updaters/test_split_scores.py), Election
    (tests/updaters/test_election.py) and the contiguous constraint
    (tests/constraints/test_validity.py:73-99), evaluated through the Partition API on the
    plans whose reference values tests/golden/stats_*.npz holds."""
    from gerrychain_b200 import Election, tally_region_splits
    from gerrychain_b200.constraints import contiguous, number_of_contiguous_parts

    z, meta, g = _stats_fixture_graph(name)
    cols = meta["region_cols"]
    updaters = {"pop": Tally("pop"), "cut_edges": cut_edges, "splits": tally_region_splits(cols),
                "election": Election("synthetic", {"A": "votes_a", "B": "votes_b"})}
    for i in range(0, len(z["plans"]), 7):
        plan = z["plans"][i]
        if len(set(plan.tolist())) < meta["k"]:
            continue
        part = Partition(g, {v: int(plan[v]) for v in g.nodes}, updaters)
        assert part["splits"] == {col: int(z["edge_splits"][i, c]) for c, col in enumerate(cols)}
        res = part["election"]
        for pi, party in enumerate(("A", "B")):
            assert res.counts(party) == tuple(int(x) for x in z["vote_counts"][i, pi])
            assert res.seats(party) == int(z["seats"][i, pi]) == res.wins(party)
        assert res.total_votes() == int(z["votes"].sum())
        assert abs(res.percent("A") + res.percent("B") - 1.0) < 1e-12
        assert contiguous(part) == bool(z["is_contiguous"][i])
        assert number_of_contiguous_parts(part) == int((z["district_pieces"][i] == 1).sum())
    with pytest.raises(ValueError):
        tally_region_splits(cols)(Partition(g, {v: int(z["plans"][0][v]) for v in g.nodes}, {"pop": Tally("pop")},
                                            use_default_updaters=False))


def test_contiguous_checks_only_the_districts_a_flip_touched():
    """constraints/contiguity.py:141-181: with a parent, only affected parts are examined."""
    from gerrychain_b200.constraints import contiguous

    grid = Grid((6, 6))
    assert contiguous(grid)
    # move the corner of district 0 away: district 0 stays connected, so does the target
    child = grid.flip({(2, 2): 3})
    touched_ok = contiguous(child)
    assert touched_ok == _contiguous(child)
    # a far-away node dropped into the middle of another district splits nothing it touches
    # but leaves an island: the island's district is "affected" and reported
    island = grid.flip({(0, 0): 3})
    assert contiguous(island) is False and _contiguous(island) is False


def test_chain_batch_statistics():
    from gerrychain_b200 import BatchedMarkovChain, Election

    z, meta, g = _stats_fixture_graph("delaunay600")
    k = meta["k"]
    plan0 = z["plans"][0]
    init = Partition(g, {v: int(plan0[v]) for v in g.nodes}, {"pop": Tally("pop"), "cut_edges": cut_edges})
    ideal = sum(init["pop"].values()) / k
    proposal = functools.partial(recom, pop_col="pop", pop_target=ideal, epsilon=0.1, node_repeats=1)
    batch = BatchedMarkovChain(proposal, [within_percent_of_ideal_population(init, 0.1, pop_key="pop")], always_accept,
                               init, total_steps=26, n_chains=64, seed=5)
    last = batch.run_to_end()
    assert last.contiguous().all() and (last.district_pieces() == 1).all()
    splits = last.region_splits("county")
    election = Election("synthetic", {"A": "votes_a", "B": "votes_b"})
    votes = last.election(election)
    assert votes.shape == (64, 2, k) and (votes.sum(axis=2) == z["votes"].sum(axis=1)[None, :]).all()
    from gerrychain_b200 import tally_region_splits

    for ch in (0, 13, 63):
        part = last.partition(ch)
        assert tally_region_splits(["county"])(part) == {"county": int(splits[ch, 0])}
        assert _contiguous(part)
        res = election(part)
        assert res.counts("A") == tuple(int(x) for x in votes[ch, 0])


def test_record_and_replay_a_recom_chain(tmp_path):
    """Record(chain, file) / Replay(graph, file, updaters) (reference
    docs/topics/reproducibility.rst:21-60): the replayed run has the recorded assignments,
    and its device-backed updaters give the recorded values."""
    from gerrychain_b200 import Record, Replay

    random.seed(12)
    grid = Grid((10, 10))
    proposal = functools.partial(recom, pop_col="population", pop_target=25, epsilon=0.2, node_repeats=1)
    chain = MarkovChain(proposal, [within_percent_of_ideal_population(grid, 0.2)], always_accept, grid, total_steps=25)
    path = str(tmp_path / "run.rcd")
    seen = [(dict(s.assignment.mapping), set(s["cut_edges"]), dict(s["population"])) for s in Record(chain, path)]
    assert len(seen) == 25
    replayed = list(Replay(grid.graph, path, updaters=grid.updaters))
    assert len(replayed) == 25
    for (mapping, cuts, pops), state in zip(seen, replayed):
        assert state.assignment.mapping == mapping
        assert set(state["cut_edges"]) == cuts
        assert dict(state["population"]) == pops


def test_recom_grows_the_engine_caps_instead_of_failing():
    """Districts of very uneven node counts: the merged pair of the two large ones exceeds the
    engine's default caps (2.4 N / k nodes).  The reference has no such limit; the host API
    takes the step again on an engine with doubled caps (RC_ERR_CAPACITY is not surfaced)."""
    g = networkx.grid_2d_graph(20, 20)
    plan = {}
    for (i, j) in g.nodes:
        d = 0 if i < 8 else 1 if i < 16 else 2 if i < 18 else 3
        plan[(i, j)] = d
        g.nodes[(i, j)]["population"] = 1 if d < 2 else 4  # 160 / 160 / 160 / 160 people
    random.seed(3)
    p = Partition(g, plan, {"population": Tally("population"), "cut_edges": cut_edges})
    proposal = functools.partial(recom, pop_col="population", pop_target=160, epsilon=0.3, node_repeats=1)
    chain = MarkovChain(proposal, [within_percent_of_ideal_population(p, 0.3)], always_accept, p, total_steps=40)
    merged_big_pair = False
    for state in chain:
        assert _contiguous(state)
        assert state["cut_edges"] == _naive_cut_edges(state)
        assert state["population"] == _naive_tally(state, "population")
        if state.flips:
            merged_big_pair = merged_big_pair or len(state.flips) > 272
    assert merged_big_pair, "the test never merged the two large districts"


def test_tally_of_a_float_column_and_polsby_popper():
    """GeographicPartition-style float attributes (areas): Tally falls back to the host sum
    (the device tallies integers), NaN is skipped with the reference's warning."""
    g = networkx.grid_2d_graph(6, 6)
    rng = np.random.default_rng(0)
    for v in g.nodes:
        g.nodes[v]["population"] = 1
        g.nodes[v]["area"] = float(rng.random()) + 0.5
    plan = {(i, j): (0 if i < 3 else 1) for (i, j) in g.nodes}
    p = Partition(g, plan, {"area": Tally("area", dtype=float), "population": Tally("population")})
    want = _naive_tally(p, "area")
    assert p["area"].keys() == want.keys()
    for d in want:
        assert abs(p["area"][d] - want[d]) < 1e-9
    a00 = g.nodes[(0, 0)]["area"]
    g2 = g.copy()  # node attributes are read once per graph (the reference freezes its graphs too)
    g2.nodes[(0, 0)]["area"] = float("nan")
    q = Partition(g2, plan, {"area": Tally("area", dtype=float)})
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = q["area"]
    assert any("ignoring nan" in str(x.message) for x in w)
    assert abs(got[0] - (want[0] - a00)) < 1e-9 and abs(got[1] - want[1]) < 1e-9
