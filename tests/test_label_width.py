"""Plans with more than 255 districts: the 16-bit-label build of the library
(librecom_b200_l16.so, include/recom_b200.h: RC_LABEL_BITS) and of the oracle.

CPU part: both libraries export the header's symbols, the 16-bit oracle equals the 8-bit one
on small plans and keeps the updater invariants on a 300-district plan.  GPU part (through
the C ABI): the 16-bit engine is bit-exact with the oracle on the existing workloads, on
recorded reference steps and on the 300-district plan, statistics kernels included."""

import ctypes
import os
import re

import numpy as np
import pytest

import _golden
from gerrychain_b200 import _abi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _naive_check(csr, k, assign, tally, cut_mask, cut_count):
    uv = csr.edge_uv
    for c in range(assign.shape[0]):
        a = assign[c].astype(np.int64)
        assert a.max() < k
        assert np.array_equal(tally[c], np.bincount(a, weights=csr.pop, minlength=k).astype(np.int64))
        naive = a[uv[:, 0]] != a[uv[:, 1]]
        bits = np.unpackbits(cut_mask[c].view(np.uint8), bitorder="little")[: csr.n_edges]
        assert np.array_equal(bits.astype(bool), naive)
        assert cut_count[c] == naive.sum()


# ------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("bits", [8, 16])
def test_both_label_builds_export_the_header(bits):
    from gerrychain_b200.build import build

    build()
    header = open(os.path.join(ROOT, "include", "recom_b200.h")).read()
    declared = set(re.findall(r"\b(rc_[a-z_]+)\s*\(", header))
    lib = ctypes.CDLL(_abi.lib_path(bits))
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _abi.load(bits)
    assert loaded.rc_label_bits() == bits and loaded.rc_abi_version() == 1


def test_label_width_follows_the_district_count():
    assert _abi.label_bits_for(2) == 8 and _abi.label_bits_for(255) == 8
    assert _abi.label_bits_for(256) == 16 and _abi.label_bits_for(32767) == 16
    with pytest.raises(ValueError):
        _abi.label_bits_for(32768)


def test_create_checks_the_district_count_per_build():
    """The 8-bit build refuses 300 districts and names the 16-bit one; the 16-bit build gets past
    the argument checks (and then fails on the missing device, here on the CPU)."""
    csr, plan, cfg = workloads.many_districts(n_chains=1)
    assert cfg.n_districts == 300 and plan.dtype == np.uint16
    g, c = _abi.graph_struct(csr), cfg.to_struct()
    h = ctypes.c_void_p()
    lib8 = _abi.load(8)
    assert lib8.rc_engine_create(ctypes.byref(g), ctypes.byref(c), 0, ctypes.byref(h)) == _abi.RC_ERR_INVALID_ARG
    assert b"librecom_b200_l16.so" in lib8.rc_last_error()
    c.allow_pair_reselection = 1
    lib16 = _abi.load(16)
    assert lib16.rc_engine_create(ctypes.byref(g), ctypes.byref(c), 0, ctypes.byref(h)) == _abi.RC_ERR_INVALID_ARG
    assert b"allow_pair_reselection" in lib16.rc_last_error()


@pytest.mark.parametrize("make", [lambda: workloads.config1(n_chains=8, seed=3),
                                  lambda: workloads.config3(n_chains=4, seed=5),
                                  lambda: workloads.config5(n_chains=4, seed=6),
                                  lambda: workloads.config3(n_chains=4, seed=7, tree_kind=_abi.RC_TREE_WILSON)],
                         ids=["config1", "config3", "config5", "config3-wilson"])
def test_oracle_with_16_bit_labels_equals_the_8_bit_oracle(make):
    from oracle import oracle

    csr, plan, cfg = make()
    a = oracle.OracleChains(csr, cfg, plan).run(15, n_threads=4)
    b = oracle.OracleChains(csr, cfg, plan, label_bits=16).run(15, n_threads=4)
    assert a.assignment.dtype == np.uint8 and b.assignment.dtype == np.uint16
    assert (a.status == 0).all() and (b.status == 0).all()
    assert np.array_equal(a.assignments(), b.assignments())
    for x in ("tally", "cut_mask", "cut_count", "counters", "proposal_no", "step_no"):
        assert np.array_equal(getattr(a, x), getattr(b, x)), x


def test_oracle_300_districts_keeps_the_updater_invariants():
    from oracle import oracle

    csr, plan, cfg = workloads.many_districts(n_chains=8, seed=7)
    o = oracle.OracleChains(csr, cfg, plan).run(300, n_threads=8)
    assert (o.status == 0).all() and (o.step_no == 300).all()
    a = o.assignments()
    assert a.dtype == np.uint16 and int(a.max()) == 299 and (a[0] != plan).mean() > 0.3
    _naive_check(csr, 300, a, o.tally, o.cut_mask, o.cut_count)
    assert (o.tally >= cfg.pop_lower).all() and (o.tally <= cfg.pop_upper).all()
    for c in range(2):
        assert (oracle.contiguity(csr, 300, a[c]) == 1).all()


def test_record_replay_with_more_than_255_districts(tmp_path):
    import networkx

    from gerrychain_b200 import Partition, Record, Replay

    g = networkx.grid_2d_graph(40, 30)
    plan = {v: (v[0] // 2) * 15 + v[1] // 2 for v in g.nodes}  # 20 x 15 blocks of 2 x 2
    first = Partition(g, plan, {})
    assert len(first) == 300 and first.assignment_vector().dtype == np.uint16
    states = [first, first.flip({(0, 0): 299, (39, 29): 256}), ]
    states.append(states[-1].flip({(5, 5): 280}))
    path = str(tmp_path / "run.rcd")
    assert list(Record(iter(states), path)) == states
    back = list(Replay(g, path, updaters={}))
    assert [b.assignment.mapping for b in back] == [s.assignment.mapping for s in states]
    assert back[1].flips == {(0, 0): 299, (39, 29): 256}


# ------------------------------------------------------------------------------ GPU
def _engine(csr, cfg, plan, bits=None):
    from gerrychain_b200.engine import Engine

    return Engine(csr, cfg, plan, label_bits=bits)


def _same_state(eng, ora, csr, k, counters=True):
    assert np.array_equal(eng.status.cpu().numpy(), ora.status) and (ora.status == 0).all()
    assert np.array_equal(eng.assignments(), ora.assignments())
    assert np.array_equal(eng.tally.cpu().numpy(), ora.tally)
    assert np.array_equal(eng.cut_count.cpu().numpy(), ora.cut_count)
    assert np.array_equal(eng.cut_mask.cpu().numpy().view(np.uint32), ora.cut_mask)
    assert np.array_equal(eng.proposal_no.cpu().numpy().view(np.uint32), ora.proposal_no)
    assert np.array_equal(eng.step_no.cpu().numpy(), ora.step_no)
    ce, co = eng.counters.cpu().numpy(), ora.counters
    for i in (_abi.RC_CTR_TREES, _abi.RC_CTR_ATTEMPTS, _abi.RC_CTR_NODES, _abi.RC_CTR_SLOTS, _abi.RC_CTR_INVALID):
        assert not counters or np.array_equal(ce[:, i], co[:, i]), i
    _naive_check(csr, k, eng.assignments(), eng.tally.cpu().numpy(), eng.cut_mask.cpu().numpy(),
                 eng.cut_count.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("make,steps", [
    (lambda: workloads.config1(n_chains=64, seed=3), 40),
    (lambda: workloads.config2(n_chains=24, seed=4), 12),
    (lambda: workloads.config3(n_chains=32, seed=5), 20),
    (lambda: workloads.config5(n_chains=16, seed=6), 15),
    (lambda: workloads.config3(n_chains=16, seed=7, tree_kind=_abi.RC_TREE_WILSON), 10),
    (lambda: workloads.config1(n_chains=24, seed=8, tree_kind=_abi.RC_TREE_WILSON,
                               proposal_kind=_abi.RC_PROPOSAL_REVERSIBLE, rev_M=30), 400),
], ids=["config1", "config2", "config3", "config5", "config3-wilson", "config1-reversible"])
def test_16_bit_build_bit_exact_with_the_oracle_on_the_named_workloads(make, steps):
    """The library built with 16-bit labels, forced onto plans that would fit a byte: the same
    chains as the (8-bit) oracle, i.e. the same chains as the default build."""
    from oracle import oracle

    csr, plan, cfg = make()
    eng = _engine(csr, cfg, plan, bits=16)
    assert eng.label_bits == 16 and eng.assignments().dtype == np.uint16
    eng.run(steps).synchronize()
    ora = oracle.OracleChains(csr, cfg, plan).run(steps, n_threads=8)
    _same_state(eng, ora, csr, cfg.n_districts, counters=cfg.proposal_kind == _abi.RC_PROPOSAL_RECOM)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["replay_grid20_mst", "replay_delaunay2400_region", "replay_grid20_wilson"])
def test_16_bit_build_replays_the_reference(name):
    g = _golden.load(name)
    eng = _engine(g.csr, g.cfg, g.init_assignment, bits=16)
    trace, info = eng.replay(g.rec)
    assert trace.dtype == np.uint16
    assert (info[:, 3] == 0).all(), [_abi.STATUS_NAMES.get(int(x)) for x in info[:, 3]]
    assert np.array_equal(trace, g.assignment)
    assert np.array_equal(info[:, 2], g.n_cut_edges)
    assert np.array_equal(info[:, 0], np.diff(g.tree_ptr))


@pytest.mark.gpu
@pytest.mark.parametrize("wilson", [False, True])
def test_300_districts_bit_exact_with_the_oracle(wilson):
    from oracle import oracle

    kw = dict(tree_kind=_abi.RC_TREE_WILSON) if wilson else {}
    csr, plan, cfg = workloads.many_districts(n_chains=96, seed=11, **kw)
    eng = _engine(csr, cfg, plan)
    assert eng.label_bits == 16
    eng.run(150).synchronize()
    ora = oracle.OracleChains(csr, cfg, plan).run(150, n_threads=8)
    _same_state(eng, ora, csr, 300)
    a = eng.assignments()
    assert int(a.max()) == 299 and (a[0] != plan).mean() > 0.2
    # the statistics kernels over 16-bit rows
    assert (eng.contiguity().cpu().numpy() == 1).all()
    col = (np.arange(csr.n_nodes) * 7 % 13).astype(np.int32)
    want = np.stack([np.bincount(a[c].astype(np.int64), weights=col, minlength=300) for c in range(a.shape[0])])
    assert np.array_equal(eng.tally_column(col).cpu().numpy(), want.astype(np.int64))
    region = (np.arange(csr.n_nodes) // 90).astype(np.int32)  # 40 bands of 90 nodes
    got = eng.region_splits(region, 40).cpu().numpy()
    for c in (0, 5, 95):
        assert tuple(int(x) for x in got[c]) == oracle.region_splits(csr, 300, a[c], region, 40)
    cut, dev = eng.histograms(4096, 64, 0.01)
    assert int(cut.sum()) == 96 and int(dev.sum()) == 96
    assert np.array_equal(np.bincount(ora.cut_count, minlength=4096), cut.cpu().numpy())


@pytest.mark.gpu
def test_300_districts_through_host_buffers_and_seed_plans():
    """rc_engine_step_host with 16-bit host rows (pitch in labels), and recursive_tree_part for
    300 districts against the oracle."""
    from dataclasses import replace

    from oracle import oracle

    csr, plan, cfg = workloads.many_districts(n_chains=32, seed=13)
    eng = _engine(csr, cfg, plan)
    h_assign, h_tally, h_cut, h_status = eng.host_buffers()
    h_assign.copy_(eng.torch.from_numpy(np.broadcast_to(plan, (32, csr.n_nodes)).copy().view(np.int16)))
    eng.step_host(h_assign, 25, h_assign, h_tally, h_cut, h_status)
    ora = oracle.OracleChains(csr, cfg, plan).run(25, n_threads=8)
    assert int(h_status.max()) == 0
    assert np.array_equal(h_assign.numpy().view(np.uint16), ora.assignments())
    assert np.array_equal(h_tally.numpy(), ora.tally) and np.array_equal(h_cut.numpy(), ora.cut_count)
    # numpy rows with their own pitch (two-dimensional copy)
    rows = np.zeros((32, csr.n_nodes + 24), dtype=np.uint16)
    rows[:, : csr.n_nodes] = plan
    view = rows[:, : csr.n_nodes]
    eng.set_assignment(plan)
    eng.proposal_no.zero_()
    eng.step_host(view, 25, view)
    assert np.array_equal(view, ora.assignments())

    scfg = replace(cfg, n_chains=4, cap_nodes=csr.n_nodes, cap_edges=csr.n_edges, seed=1)
    seng = _engine(csr, scfg, np.zeros(csr.n_nodes, dtype=np.uint16))
    seng.seed(10000).synchronize()
    status, plans, tally, _ = oracle.seed_plans(csr, scfg, n_plans=4)
    assert np.array_equal(seng.status.cpu().numpy(), status) and (status == 0).all()
    assert np.array_equal(seng.assignments(), plans)
    assert np.array_equal(seng.tally.cpu().numpy(), tally)
    assert (seng.contiguity().cpu().numpy() == 1).all()


@pytest.mark.gpu
def test_batched_chain_of_a_300_district_partition():
    import functools

    import networkx

    from oracle import oracle
    from gerrychain_b200 import BatchedMarkovChain, Partition
    from gerrychain_b200.accept import always_accept
    from gerrychain_b200.config import EngineConfig
    from gerrychain_b200.constraints import within_percent_of_ideal_population
    from gerrychain_b200.proposals import recom
    from gerrychain_b200.updaters import Tally, cut_edges

    g = networkx.grid_2d_graph(60, 60)
    networkx.set_node_attributes(g, 1, "population")
    plan = {v: (v[0] // 3) * 15 + v[1] // 4 for v in g.nodes}
    init = Partition(g, plan, {"population": Tally("population"), "cut_edges": cut_edges})
    assert len(init) == 300
    p = functools.partial(recom, pop_col="population", pop_target=12.0, epsilon=0.1)
    cons = [within_percent_of_ideal_population(init, 0.1)]
    runner = BatchedMarkovChain(p, cons, always_accept, init, total_steps=31, n_chains=16, seed=77)
    last = runner.run_to_end()
    eng = runner.engine
    assert eng.label_bits == 16
    cfg = EngineConfig(n_districts=300, n_chains=16, pop_target=12.0, epsilon=0.1, seed=77,
                       pop_lower=cons[0].bounds[0], pop_upper=cons[0].bounds[1])
    ora = oracle.OracleChains(eng.csr, cfg, init.assignment_vector()).run(30, n_threads=4)
    assert np.array_equal(last.assignment.cpu().numpy().view(np.uint16), ora.assignments())
    part = last.partition(3)
    assert len(part) == 300 and len(part["cut_edges"]) == int(ora.cut_count[3])
    assert sorted(part["population"].values()) == sorted(int(x) for x in ora.tally[3])


@pytest.mark.gpu
def test_single_chain_recom_with_300_districts():
    """The literal drop-in (MarkovChain + recom, host Partition per state) on a 300-district plan:
    incremental cut_edges / Tally equal the naive recomputation at every step, the child keeps the
    parent's uint16 dense plan, districts stay contiguous."""
    import functools

    import networkx

    from gerrychain_b200 import MarkovChain, Partition
    from gerrychain_b200.accept import always_accept
    from gerrychain_b200.constraints import within_percent_of_ideal_population
    from gerrychain_b200.proposals import recom
    from gerrychain_b200.updaters import Tally, cut_edges

    g = networkx.grid_2d_graph(60, 60)
    networkx.set_node_attributes(g, 1, "population")
    plan = {v: (v[0] // 3) * 15 + v[1] // 4 for v in g.nodes}
    init = Partition(g, plan, {"population": Tally("population"), "cut_edges": cut_edges})
    p = functools.partial(recom, pop_col="population", pop_target=12.0, epsilon=0.1)
    chain = MarkovChain(p, [within_percent_of_ideal_population(init, 0.1)], always_accept, init, total_steps=40)
    moved = 0
    for state in chain:
        assert len(state) == 300 and state.assignment_vector().dtype == np.uint16
        m = state.assignment.mapping
        assert state["cut_edges"] == {tuple(sorted(e)) for e in g.edges if m[e[0]] != m[e[1]]}
        pops = {}
        for node, part in m.items():
            pops[part] = pops.get(part, 0) + 1
        assert state["population"] == pops and set(pops.values()) <= {11, 12, 13}
        if state.parent is not None:
            moved += len(state.flips) > 0
            for d in set(state.flips.values()):  # the districts that gained nodes
                assert networkx.is_connected(g.subgraph(state.assignment.parts[d]))
    assert moved >= 10
