"""CPU tests of the host-side mirror of GerryChain's interface (no GPU, no engine):
MarkovChain's iterator contract (reference tests/test_chain.py:23-88), Validator / Bounds
(reference tests/constraints/test_bounds.py:5-43, test_validity.py), recom-argument parsing,
chain sharding and the histogram all-reduce over gloo (world_size 2)."""

import functools
import os
import subprocess
import sys
import textwrap
from unittest.mock import MagicMock

import pytest

import gerrychain_b200 as gc
from gerrychain_b200 import _abi
from gerrychain_b200.chain import MarkovChain, _device_bounds, _recom_arguments
from gerrychain_b200.constraints import Bounds, Validator, within_percent_of_ideal_population
from gerrychain_b200.distributed import shard_chains
from gerrychain_b200.proposals import engine_kwargs, method_config, raise_for_status

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class MockState:
    def flip(self, changes):
        return MockState()


def mock_proposal(state):
    return state.flip({1: 2})


def mock_accept(state):
    return True


def mock_is_valid(state):
    return True


def test_markov_chain_yields_total_steps_and_first_is_initial():
    initial = MockState()
    chain = MarkovChain(mock_proposal, mock_is_valid, mock_accept, initial, total_steps=10)
    states = list(chain)
    assert len(states) == 10
    assert states[0] is initial
    assert len(chain) == 10
    assert repr(chain) == "<MarkovChain [10 steps]>"
    assert len(list(chain)) == 10  # __iter__ resets


def test_markov_chain_rejects_invalid_initial_state():
    def never(state):
        return False

    with pytest.raises(ValueError) as e:
        MarkovChain(mock_proposal, [never], mock_accept, MockState(), total_steps=10)
    assert "never" in str(e.value)


def test_markov_chain_only_counts_valid_and_reyields_rejected():
    calls = {"n": 0}

    def valid_every_other(state):
        calls["n"] += 1
        return calls["n"] % 2 == 1

    initial = MockState()
    rejecting = MarkovChain(mock_proposal, valid_every_other, lambda s: False, initial, total_steps=5)
    states = list(rejecting)
    assert len(states) == 5
    assert all(s is initial for s in states)  # accept() False: the old state is yielded again


def test_validator_needs_real_booleans():
    v = Validator([lambda p: 1])
    with pytest.raises(TypeError):
        v(MagicMock())
    import numpy

    assert Validator([lambda p: numpy.bool_(True)])(MagicMock()) is True
    assert Validator([lambda p: True, lambda p: False])(MagicMock()) is False


def test_within_percent_of_ideal_population_on_plain_dicts():
    # reference tests/constraints/test_bounds.py: a dict works as the "partition"
    initial = {"population": {0: 100, 1: 100, 2: 100}}
    c = within_percent_of_ideal_population(initial, percent=0.05)
    assert isinstance(c, Bounds) and c.pop_key == "population"
    assert c.bounds == (95.0, 105.0)
    assert c({"population": {0: 95, 1: 105, 2: 100}}) is True
    assert c({"population": {0: 94, 1: 106, 2: 100}}) is False
    assert c.__name__.startswith("Bounds(population")


def test_method_config_reads_partials_like_the_reference():
    assert method_config(gc.bipartition_tree)["tree_kind"] == _abi.RC_TREE_MST
    m = functools.partial(gc.bipartition_tree, max_attempts=7, allow_pair_reselection=True,
                          spanning_tree_fn=gc.tree.uniform_spanning_tree)
    cfg = method_config(functools.partial(m, warn_attempts=3))
    assert cfg == dict(tree_kind=_abi.RC_TREE_WILSON, cut="region", max_attempts=7, warn_attempts=3,
                       allow_pair_reselection=True)
    with pytest.raises(NotImplementedError):
        method_config(functools.partial(gc.bipartition_tree, choice=min))
    import random

    # the reference's defaults spelled out, and the contraction variant of the balance test
    assert method_config(functools.partial(gc.bipartition_tree, choice=random.choice,
                                           balance_edge_fn=gc.tree.find_balanced_edge_cuts_contraction)) \
        == method_config(gc.bipartition_tree)
    with pytest.raises(NotImplementedError):
        method_config(functools.partial(gc.bipartition_tree, balance_edge_fn=len))
    with pytest.raises(NotImplementedError):
        method_config(lambda *a, **k: None)
    kw = engine_kwargs(100, 0.1, 2, {"county": 0.5}, gc.bipartition_tree)
    assert kw["cut_policy"] == _abi.RC_CUT_REGION_PREF and kw["node_repeats"] == 2
    kw = engine_kwargs(100, 0.1, 1, {"a": 1, "b": 1, "c": 1, "d": 1}, gc.bipartition_tree)
    assert kw["cut_policy"] == _abi.RC_CUT_MAX_WEIGHT  # tree.py:548: more than 3 regions


def test_selectors_have_no_host_implementation():
    with pytest.raises(NotImplementedError):
        gc.tree.random_spanning_tree(None)
    with pytest.raises(NotImplementedError):
        gc.bipartition_tree(None, "p", 1, 0.1)


def test_status_words_map_to_the_reference_exceptions():
    with pytest.raises(RuntimeError, match="Could not find a possible cut after 1 attempts."):
        raise_for_status(_abi.RC_ERR_MAX_ATTEMPTS, 1)  # reference tests/test_region_aware.py:102-112
    with pytest.raises(gc.MetagraphError):
        raise_for_status(_abi.RC_ERR_ALL_PAIRS_BAD, 10, 6.0)
    with pytest.raises(IndexError):
        raise_for_status(_abi.RC_ERR_NO_CUT_EDGES, 10)
    raise_for_status(_abi.RC_OK, 10)


def test_batched_runner_refuses_host_only_callables():
    p = functools.partial(gc.recom, pop_col="population", pop_target=10, epsilon=0.1)
    assert _recom_arguments(p)[:4] == ("population", 10, 0.1, 1)
    assert _recom_arguments(gc.ReCom("population", 10, 0.1))[0] == "population"
    with pytest.raises(TypeError):
        _recom_arguments(lambda s: s)
    with pytest.raises(TypeError):
        _device_bounds([lambda p: True], "population")
    b = within_percent_of_ideal_population({"population": {0: 10, 1: 10}}, 0.1)
    _, lo, hi = _device_bounds(Validator([b]), "population")
    assert (lo, hi) == (9.0, 11.0)


def test_shard_chains_covers_every_chain_once():
    for total, world in ((8192, 8), (1000, 3), (5, 8)):
        spans = [shard_chains(total, r, world) for r in range(world)]
        assert spans[0][0] == 0
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1
        assert spans[-1][0] + spans[-1][1] == total


def test_histogram_all_reduce_over_gloo_world_size_2(tmp_path):
    """The only collective of the path, exercised on CPU tensors with two processes."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from gerrychain_b200.distributed import all_reduce_histograms, gather_chain_scalars, shard_chains
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        off, n = shard_chains(10, r, w)
        cut = torch.zeros(8, dtype=torch.int64); cut[r] = n
        dev = torch.ones(4, dtype=torch.int64) * (r + 1)
        all_reduce_histograms(cut, dev)
        assert cut.tolist()[:2] == [5, 5] and dev.tolist() == [3] * 4, (cut, dev)
        vals = gather_chain_scalars(torch.arange(off, off + n, dtype=torch.int32))
        assert vals.tolist() == list(range(10)), vals
        dist.destroy_process_group()
        print("rank%d_ok" % r, flush=True)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "rank0_ok" in out.stdout and "rank1_ok" in out.stdout


def test_graph_json_round_trip_feeds_the_csr_builder(tmp_path):
    """Graph.from_json / to_json (reference graph/graph.py:76-119): networkx adjacency JSON
    -> Graph -> canonical CSR, node order and attributes preserved."""
    import networkx
    import numpy as np

    from gerrychain_b200.csr import csr_from_networkx

    g = networkx.grid_2d_graph(3, 4)
    g = networkx.relabel_nodes(g, {n: f"v{n[0]}{n[1]}" for n in g.nodes})
    for i, n in enumerate(g.nodes):
        g.nodes[n].update(TOTPOP=10 + i, county="A" if i < 6 else None)
    path = tmp_path / "dual.json"
    gc.Graph.from_networkx(g).to_json(str(path))
    loaded = gc.Graph.from_json(str(path))
    assert repr(loaded) == "<Graph [12 nodes, 17 edges]>"
    assert list(loaded.nodes) == list(g.nodes) and loaded.lookup("v00", "TOTPOP") == 10
    a = csr_from_networkx(g, "TOTPOP", {"county": 0.5})
    b = csr_from_networkx(loaded, "TOTPOP", {"county": 0.5})
    for f in ("row_ptr", "col_idx", "slot_edge", "edge_uv", "pop", "region_ids"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert (b.region_ids[0] == -1).sum() == 6  # None -> -1: always surcharged (tree.py:84-85)
    island = networkx.Graph()
    island.add_nodes_from([0, 1, 2])
    island.add_edge(0, 1)
    with pytest.warns(UserWarning, match="islands"):
        gc.Graph.from_networkx(island).issue_warnings()


def test_record_replay_round_trip(tmp_path):
    """Record / Replay (reference docs/topics/reproducibility.rst:21-60): every recorded state
    comes back with the same assignment, parts and flips; the file is a fraction of the
    dense plans.  Host logic only - no device value is touched."""
    import random

    import networkx

    from gerrychain_b200 import Partition, Record, Replay

    g = networkx.grid_2d_graph(12, 12)
    rng = random.Random(3)
    labels = ["north", "south", "east", "west"]
    plan = {v: labels[(v[0] >= 6) * 2 + (v[1] >= 6)] for v in g.nodes}
    states = [Partition(g, plan, {})]
    for _ in range(40):
        flips = {rng.choice(list(g.nodes)): rng.choice(labels) for _ in range(rng.randint(0, 30))}
        states.append(states[-1].flip(flips))
    path = str(tmp_path / "run.rcd")
    passed = list(Record(iter(states), path))
    assert passed == states  # passed through untouched
    replayed = list(Replay(g, path, updaters={}))
    assert len(replayed) == len(states)
    for a, b in zip(states, replayed):
        assert a.assignment.mapping == b.assignment.mapping
        assert a.parts == b.parts
        assert isinstance(b, Partition) and len(b) == 4
    for prev, cur in zip(replayed, replayed[1:]):
        changed = {v: cur.assignment.mapping[v] for v in g.nodes if cur.assignment.mapping[v] != prev.assignment.mapping[v]}
        assert cur.flips == changed
    import os

    assert os.path.getsize(path) < 0.2 * len(states) * 144
    with pytest.raises(ValueError):
        list(Replay(networkx.grid_2d_graph(3, 3), path))
    bad = tmp_path / "bad.rcd"
    bad.write_bytes(b"nope")
    with pytest.raises(ValueError):
        list(Replay(g, str(bad)))


def test_partition_from_districtr_file(tmp_path):
    """partition/partition.py:253-296: the plan is keyed by the unit id column."""
    import json

    import networkx as nx

    from gerrychain_b200 import Partition

    g = nx.path_graph(4)
    for v in g.nodes:
        g.nodes[v]["GEOID"] = 100 + v
    path = tmp_path / "plan.json"
    path.write_text(json.dumps({"idColumn": {"key": "GEOID"}, "assignment": {"100": 0, "101": 0, "102": 1, "103": 1}}))
    p = Partition.from_districtr_file(g, str(path))
    assert dict(p.assignment.items()) == {0: 0, 1: 0, 2: 1, 3: 1} and len(p) == 2
    path.write_text(json.dumps({"idColumn": {"key": "missing"}, "assignment": {}}))
    with pytest.raises(TypeError):
        Partition.from_districtr_file(g, str(path))


def test_subgraphs_and_num_spanning_trees():
    """partition/subgraphs.py and updaters/spanning_trees.py: a 3 x 3 block of a lattice has
    192 spanning trees, a 2 x 2 block 4, a path 1."""
    import networkx as nx

    from gerrychain_b200 import Partition
    from gerrychain_b200.updaters import num_spanning_trees

    g = nx.grid_2d_graph(3, 6)
    plan = {(i, j): 0 if j < 3 else (1 if i < 2 and j < 5 else 2) for (i, j) in g.nodes}
    p = Partition(g, plan, {"trees": num_spanning_trees})
    assert set(p.subgraphs[1].nodes) == {(0, 3), (0, 4), (1, 3), (1, 4)}
    assert sorted(len(sub) for sub in p.subgraphs) == [4, 5, 9]
    counts = p["trees"]
    assert round(counts[0]) == 192 and round(counts[1]) == 4 and round(counts[2]) == 1
    assert abs(counts[0] - 192) < 1e-9


def test_batched_runner_accepts_population_bounds_and_contiguous_only():
    """chain.py (batched runner): which constraints the device enforces."""
    from gerrychain_b200 import constraints as C
    from gerrychain_b200.chain import _device_bounds

    b1 = C.Bounds(lambda p: p["population"].values(), (90.0, 110.0)); b1.pop_key = "population"
    b2 = C.Bounds(lambda p: p["population"].values(), (95.0, 120.0)); b2.pop_key = "population"
    cons, lo, hi = _device_bounds([b1, C.contiguous, b2, C.single_flip_contiguous], "population")
    assert (lo, hi) == (95.0, 110.0) and len(cons) == 4
    cons, lo, hi = _device_bounds(C.Validator([C.contiguous]), "population")
    assert lo == -float("inf") and hi == float("inf")
    with pytest.raises(TypeError):
        _device_bounds([C.UpperBound(lambda p: 1, 2)], "population")
    with pytest.raises(TypeError):
        _device_bounds([C.Bounds(lambda p: [1], (0, 2))], "population")  # not a population bound


def test_data_tally_over_external_values():
    """updaters/tally.py:12-66: values from a mapping, a Series or an attribute name; NaN
    is skipped with a warning."""
    import networkx as nx
    import pandas

    from gerrychain_b200 import Partition
    from gerrychain_b200.updaters import DataTally

    g = nx.path_graph(6)
    for v in g.nodes:
        g.nodes[v]["w"] = v * 0.5
    plan = {v: v // 3 for v in g.nodes}
    p = Partition(g, plan, {"ints": DataTally({v: v + 1 for v in g.nodes}, "ints"),
                            "series": DataTally(pandas.Series({v: float(v) for v in g.nodes}), "series"),
                            "attr": DataTally("w", "attr")})
    assert p["ints"] == {0: 6, 1: 15} and all(isinstance(x, int) for x in p["ints"].values())
    assert p["series"] == {0: 3.0, 1: 12.0}
    assert p["attr"] == {0: 1.5, 1: 6.0}
    assert p.flip({2: 1})["ints"] == {0: 3, 1: 18}
    holes = Partition(g, plan, {"d": DataTally({v: (float("nan") if v == 4 else 1.0) for v in g.nodes}, "d")})
    with pytest.warns(UserWarning, match="ignoring nan"):
        assert holes["d"] == {0: 3.0, 1: 2.0}
