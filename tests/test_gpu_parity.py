"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI,
against (1) recorded reference steps and (2) the CPU oracle on identical Philox draws."""

import numpy as np
import pytest

import _golden
from gerrychain_b200 import _abi, workloads

pytestmark = pytest.mark.gpu


def _engine(csr, cfg, plan):
    from gerrychain_b200.engine import Engine

    return Engine(csr, cfg, plan)


def _naive_check(csr, k, assign, tally, cut_mask, cut_count):
    """Incremental updaters == naive recomputation (reference tests/updaters/
    test_cut_edges.py:144-161, tests/test_tally.py:56-76)."""
    uv = csr.edge_uv
    for c in range(assign.shape[0]):
        a = assign[c]
        want = np.bincount(a, weights=csr.pop, minlength=k).astype(np.int64)
        assert np.array_equal(tally[c], want)
        naive = a[uv[:, 0]] != a[uv[:, 1]]
        bits = np.unpackbits(cut_mask[c].view(np.uint8), bitorder="little")[: csr.n_edges]
        assert np.array_equal(bits.astype(bool), naive)
        assert cut_count[c] == naive.sum()


@pytest.mark.parametrize("name", _golden.names("mst"))
def test_replay_bit_exact_with_reference(name):
    g = _golden.load(name)
    eng = _engine(g.csr, g.cfg, g.init_assignment)
    trace, info = eng.replay(g.rec)
    assert (info[:, 3] == 0).all(), [_abi.STATUS_NAMES.get(int(x)) for x in info[:, 3]]
    assert np.array_equal(trace, g.assignment)
    assert np.array_equal(info[:, 2], g.n_cut_edges)
    assert np.array_equal(info[:, 1], np.diff(g.balanced_ptr))
    assert np.array_equal(info[:, 0], np.diff(g.tree_ptr))
    assert np.array_equal(eng.tally.cpu().numpy()[0], g.population[-1])
    _naive_check(g.csr, g.cfg.n_districts, eng.assignments(), eng.tally.cpu().numpy(),
                 eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())


@pytest.mark.parametrize("walkers", [0, 1])
@pytest.mark.parametrize("name", _golden.names("wilson"))
def test_wilson_replay_bit_exact_with_reference(name, walkers):
    """uniform_spanning_tree (tree.py:97-132): the reference's recorded per-node successor
    draws, replayed through the device's lane walkers (the recorded trees of a step walked
    concurrently, or one at a time), give the reference's assignment after every step."""
    g = _golden.load(name)
    g.cfg.wilson_walkers = walkers
    eng = _engine(g.csr, g.cfg, g.init_assignment)
    trace, info = eng.replay(g.rec)
    assert (info[:, 3] == 0).all(), [_abi.STATUS_NAMES.get(int(x)) for x in info[:, 3]]
    assert np.array_equal(trace, g.assignment)
    assert np.array_equal(info[:, 2], g.n_cut_edges)
    assert np.array_equal(info[:, 1], np.diff(g.balanced_ptr))
    assert np.array_equal(info[:, 0], np.diff(g.tree_ptr))
    # every recorded draw was consumed: same number of walk steps as the reference
    assert int(eng.counters.cpu().numpy()[0, _abi.RC_CTR_WALK]) == len(g.rec.stack_val)
    _naive_check(g.csr, g.cfg.n_districts, eng.assignments(), eng.tally.cpu().numpy(),
                 eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())


@pytest.mark.parametrize("walkers", [0, 1, 7])
@pytest.mark.parametrize("name", [n for n in _golden.names("wilson") if "200000" not in n])
def test_wilson_free_run_bit_exact_with_oracle(name, walkers):
    """walkers: how many trees of one bipartition_tree call are walked concurrently - the
    state after every step, the tree counts and the per-tree draw counts are the same."""
    g = _golden.load(name)
    cfg = g.cfg
    cfg.n_chains = 16
    cfg.seed = 4321
    cfg.wilson_walkers = walkers
    eng = _free_run_vs_oracle(g.csr, g.init_assignment, cfg, 30)
    from oracle import oracle  # walk-step counts agree as well

    ora = oracle.OracleChains(g.csr, cfg, g.init_assignment).run(30, n_threads=8)
    assert np.array_equal(eng.counters.cpu().numpy()[:, _abi.RC_CTR_WALK], ora.counters[:, _abi.RC_CTR_WALK])


@pytest.mark.parametrize("walkers", [0, 2])
def test_wilson_free_run_config3(walkers):
    csr, plan, cfg = workloads.config3(n_chains=16, seed=21, tree_kind=_abi.RC_TREE_WILSON,
                                       wilson_walkers=walkers)
    _free_run_vs_oracle(csr, plan, cfg, 10)


@pytest.mark.parametrize("name", [n for n in _golden.names("mst") if "region" in n])
def test_replay_region_policy_on_device(name):
    g = _golden.load(name, inject_cut=False)
    eng = _engine(g.csr, g.cfg, g.init_assignment)
    trace, info = eng.replay(g.rec)
    assert (info[:, 3] == 0).all()
    assert np.array_equal(trace, g.assignment)


def _free_run_vs_oracle(csr, plan, cfg, steps):
    from oracle import oracle

    eng = _engine(csr, cfg, plan)
    eng.run(steps).synchronize()
    ora = oracle.OracleChains(csr, cfg, plan).run(steps, n_threads=8)
    assert np.array_equal(eng.status.cpu().numpy(), ora.status)
    assert (ora.status == 0).all()
    assert np.array_equal(eng.assignments(), ora.assignments())
    assert np.array_equal(eng.tally.cpu().numpy(), ora.tally)
    assert np.array_equal(eng.cut_count.cpu().numpy(), ora.cut_count)
    assert np.array_equal(eng.cut_mask.cpu().numpy().view(np.uint32), ora.cut_mask)
    assert np.array_equal(eng.proposal_no.cpu().numpy().view(np.uint32), ora.proposal_no)
    assert np.array_equal(eng.step_no.cpu().numpy(), ora.step_no)
    ce, co = eng.counters.cpu().numpy(), ora.counters
    for i in (_abi.RC_CTR_TREES, _abi.RC_CTR_ATTEMPTS, _abi.RC_CTR_NODES, _abi.RC_CTR_SLOTS,
              _abi.RC_CTR_INVALID):
        assert np.array_equal(ce[:, i], co[:, i]), i
    _naive_check(csr, cfg.n_districts, eng.assignments(), eng.tally.cpu().numpy(),
                 eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())
    return eng


@pytest.mark.parametrize("name", ["replay_grid20_mst", "replay_grid10diag_mst",
                                  "replay_delaunay2400_mst", "replay_delaunay2400_region",
                                  "replay_8x8_region2", "replay_8x8_region3"])
def test_free_run_bit_exact_with_oracle_small(name):
    g = _golden.load(name)
    cfg = g.cfg
    cfg.n_chains = 16
    cfg.seed = 1234
    _free_run_vs_oracle(g.csr, g.init_assignment, cfg, 40)


def test_free_run_config1():
    csr, plan, cfg = workloads.config1(n_chains=64, seed=7)
    _free_run_vs_oracle(csr, plan, cfg, 100)


def test_free_run_config2():
    csr, plan, cfg = workloads.config2(n_chains=32, seed=11)
    _free_run_vs_oracle(csr, plan, cfg, 25)


def test_free_run_config3():
    csr, plan, cfg = workloads.config3(n_chains=32, seed=13)
    _free_run_vs_oracle(csr, plan, cfg, 25)


def test_free_run_config5_region_aware():
    csr, plan, cfg = workloads.config5(n_chains=16, seed=17)
    _free_run_vs_oracle(csr, plan, cfg, 15)


def test_free_run_config2_full_batch_many_steps():
    """BASELINE's full batch (1024 chains, so that chains migrate between CTAs through the
    ready ring and every SM works on several chains) for 40 steps, in several launches of
    different length: still bit-exact with the sequential oracle."""
    from oracle import oracle

    csr, plan, cfg = workloads.config2(n_chains=1024, seed=2024)
    eng = _engine(csr, cfg, plan)
    for n in (1, 7, 32):
        eng.run(n)
    eng.synchronize()
    ora = oracle.OracleChains(csr, cfg, plan).run(40, n_threads=16)
    assert (ora.status == 0).all()
    assert np.array_equal(eng.assignments(), ora.assignments())
    assert np.array_equal(eng.tally.cpu().numpy(), ora.tally)
    assert np.array_equal(eng.cut_mask.cpu().numpy().view(np.uint32), ora.cut_mask)
    assert np.array_equal(eng.proposal_no.cpu().numpy().view(np.uint32), ora.proposal_no)
    assert np.array_equal(eng.counters.cpu().numpy()[:, :5], ora.counters[:, :5])


def test_node_repeats_and_tight_bounds():
    # percent != epsilon: proposals can be invalid and must not count as steps
    csr, plan, cfg = workloads.config1(n_chains=16, seed=3, node_repeats=2)
    ideal = cfg.pop_target
    cfg.pop_lower, cfg.pop_upper = 0.98 * ideal, 1.02 * ideal
    eng = _free_run_vs_oracle(csr, plan, cfg, 50)
    assert eng.counters.cpu().numpy()[:, _abi.RC_CTR_INVALID].sum() > 0


def test_shared_retries_bit_exact_with_oracle(monkeypatch):
    """Shared retries: CTAs with no chain to step draw the later spanning trees of a step
    another CTA is retrying (tree.py:679-718); the lowest tree with a balanced cut wins, so the
    chains - and their tree / attempt / draw counters - must not change.  Forced on for every
    launch (RECOM_B200_SHARE=1): a full grid with a tail (config 3, 640 chains, heavy-tailed
    retries), a status window of 3 units instead of 256 (RECOM_B200_HELP_WIN: every shared step
    slides it), a handful of chains with hundreds of helpers, and Wilson batches."""
    monkeypatch.setenv("RECOM_B200_SHARE", "1")
    csr, plan, cfg = workloads.config3(n_chains=640, seed=77)
    _free_run_vs_oracle(csr, plan, cfg, 12)
    monkeypatch.setenv("RECOM_B200_HELP_WIN", "3")
    csr, plan, cfg = workloads.config3(n_chains=24, seed=9)
    eng = _free_run_vs_oracle(csr, plan, cfg, 30)
    assert eng.counters.cpu().numpy()[:, _abi.RC_CTR_TREES].max() > 150  # (many steps went beyond the common path)
    monkeypatch.delenv("RECOM_B200_HELP_WIN")
    csr, plan, cfg = workloads.config3(n_chains=3, seed=78)
    _free_run_vs_oracle(csr, plan, cfg, 40)
    csr, plan, cfg = workloads.config3(n_chains=40, seed=79, tree_kind=_abi.RC_TREE_WILSON)
    eng = _free_run_vs_oracle(csr, plan, cfg, 12)
    from oracle import oracle

    ora = oracle.OracleChains(csr, cfg, plan).run(12, n_threads=8)
    assert np.array_equal(eng.counters.cpu().numpy()[:, _abi.RC_CTR_WALK], ora.counters[:, _abi.RC_CTR_WALK])


def test_shared_retries_recruiting_and_limits(monkeypatch):
    """The rest of the shared-retry protocol.  (1) Recruiting: with the thresholds turned down
    (steps leave the common path after 4 failed trees and publish after 12 although every CTA has
    a chain of its own) busy CTAs leave their chains for other chains' steps all the time, and
    lagging chains keep their CTAs - the chains must not change.  (2) max_attempts in the long
    path: a step whose trees run out ends in RC_ERR_MAX_ATTEMPTS exactly where the oracle's does
    (tree.py:718), with the same attempt counts."""
    from oracle import oracle

    monkeypatch.setenv("RECOM_B200_SHARE", "1")
    monkeypatch.setenv("RECOM_B200_HELP_AFTER", "4")
    monkeypatch.setenv("RECOM_B200_RECRUIT_AFTER", "12")
    csr, plan, cfg = workloads.config3(n_chains=1024, seed=80)
    _free_run_vs_oracle(csr, plan, cfg, 8)
    monkeypatch.setenv("RECOM_B200_UNITS_PER_RECRUIT", "4")
    csr, plan, cfg = workloads.config3(n_chains=700, seed=82, tree_kind=_abi.RC_TREE_WILSON)
    _free_run_vs_oracle(csr, plan, cfg, 3)
    csr, plan, cfg = workloads.config3(n_chains=64, seed=81, max_attempts=20)
    eng = _engine(csr, cfg, plan)
    eng.run(40).synchronize()
    ora = oracle.OracleChains(csr, cfg, plan).run(40, n_threads=8)
    st = eng.status.cpu().numpy()
    assert np.array_equal(st, ora.status)
    assert (st == _abi.RC_ERR_MAX_ATTEMPTS).any() and (st == 0).any()
    assert np.array_equal(eng.step_no.cpu().numpy(), ora.step_no)
    assert np.array_equal(eng.assignments(), ora.assignments())
    ce, co = eng.counters.cpu().numpy(), ora.counters
    for i in (_abi.RC_CTR_TREES, _abi.RC_CTR_ATTEMPTS):
        assert np.array_equal(ce[:, i], co[:, i]), i


def test_max_attempts_status():
    # reference tests/test_region_aware.py:102-112: max_attempts=1 -> RuntimeError
    csr, plan, cfg = workloads.config1(n_chains=8, seed=5, max_attempts=1)
    from oracle import oracle
    from gerrychain_b200.engine import EngineError

    eng = _engine(csr, cfg, plan)
    eng.run(50).synchronize()
    ora = oracle.OracleChains(csr, cfg, plan).run(50)
    st = eng.status.cpu().numpy()
    assert np.array_equal(st, ora.status)
    assert (st == _abi.RC_ERR_MAX_ATTEMPTS).any()
    assert np.array_equal(eng.step_no.cpu().numpy(), ora.step_no)
    with pytest.raises(EngineError):
        eng.check_status()


def test_cut_edge_accept_bit_exact_with_oracle():
    """accept.py:19-35 on the device: rejected proposals leave the state alone and count."""
    csr, plan, cfg = workloads.config1(n_chains=32, seed=6, accept_rule=_abi.RC_ACCEPT_CUT_EDGE)
    eng = _free_run_vs_oracle(csr, plan, cfg, 80)
    # some proposals were rejected: fewer distinct states than steps on at least one chain
    csr, plan, cfg0 = workloads.config1(n_chains=32, seed=6)
    eng0 = _engine(csr, cfg0, plan)
    eng0.run(80).synchronize()
    assert not np.array_equal(eng.assignments(), eng0.assignments())


def test_reversible_recom_bit_exact_with_oracle():
    """tree_proposals.py:149-267: ordered pair, one Wilson tree, bound M, seam-length acceptance."""
    for M in (30, 3):  # M = 3 also meets trees with more than M balanced cuts -> RC_ERR_REVERSIBILITY
        csr, plan, cfg = workloads.config1(n_chains=24, seed=9, tree_kind=_abi.RC_TREE_WILSON,
                                           proposal_kind=_abi.RC_PROPOSAL_REVERSIBLE, rev_M=M)
        from oracle import oracle

        eng = _engine(csr, cfg, plan)
        eng.run(400).synchronize()
        ora = oracle.OracleChains(csr, cfg, plan).run(400, n_threads=8)
        assert np.array_equal(eng.status.cpu().numpy(), ora.status)
        assert np.array_equal(eng.assignments(), ora.assignments())
        assert np.array_equal(eng.tally.cpu().numpy(), ora.tally)
        assert np.array_equal(eng.cut_count.cpu().numpy(), ora.cut_count)
        assert np.array_equal(eng.step_no.cpu().numpy(), ora.step_no)
        _naive_check(csr, cfg.n_districts, eng.assignments(), eng.tally.cpu().numpy(),
                     eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())
        if M == 30:
            assert (ora.status == 0).all() and (eng.assignments() != plan).any()
        else:
            assert (ora.status == _abi.RC_ERR_REVERSIBILITY).any()


def test_pair_reselection():
    csr, plan, cfg = workloads.config1(n_chains=8, seed=5, max_attempts=2, allow_pair_reselection=True)
    _free_run_vs_oracle(csr, plan, cfg, 60)


def test_full_size_properties_config2():
    """1024 chains at BASELINE's size: size-independent properties (naive == incremental,
    population bounds, contiguity of every district)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    csr, plan, cfg = workloads.config2(n_chains=1024, seed=99)
    eng = _engine(csr, cfg, plan)
    eng.run(10).synchronize()
    eng.check_status()
    a = eng.assignments()
    tally = eng.tally.cpu().numpy()
    _naive_check(csr, cfg.n_districts, a, tally, eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())
    assert (tally >= cfg.pop_lower).all() and (tally <= cfg.pop_upper).all()
    assert (eng.step_no.cpu().numpy() == 10).all()
    uv = csr.edge_uv
    for c in range(0, 1024, 64):
        same = a[c][uv[:, 0]] == a[c][uv[:, 1]]
        m = coo_matrix((np.ones(same.sum()), (uv[same, 0], uv[same, 1])), shape=(csr.n_nodes,) * 2)
        ncomp, _ = connected_components(m, directed=False)
        assert ncomp == cfg.n_districts
    # different chains diverge (distinct Philox subsequences)
    assert len({a[c].tobytes() for c in range(64)}) == 64


def test_spill_mode_bit_exact_with_oracle():
    """Caps too large for shared memory push the working arrays into the per-CTA global
    scratch (the path 200k-node graphs take); results must not change."""
    for tree_kind, cap_nodes in ((_abi.RC_TREE_MST, 9000), (_abi.RC_TREE_WILSON, 5400), (_abi.RC_TREE_WILSON, 9000)):
        csr, plan, cfg = workloads.config3(n_chains=24, seed=31, tree_kind=tree_kind,
                                           cap_nodes=cap_nodes, cap_edges=26976)
        eng = _free_run_vs_oracle(csr, plan, cfg, 12)
        if tree_kind == _abi.RC_TREE_MST:
            assert eng.info().smem_bytes < 16 * 1024  # only the membership bitmap stays on chip
        elif cap_nodes == 5400:
            assert eng.info().spill == 1  # walker state on chip (3 CTAs per SM), the adjacency in the CTA's L2 block
        else:
            assert eng.info().spill == 2  # level 1 would leave one CTA per SM: everything in the global block, 4 CTAs


def test_wilson_all_in_l2_bit_exact_with_oracle():
    """Merged pairs of 30k nodes (a 240 x 250 grid in quadrants): neither the adjacency nor the
    walkers' state fits shared memory (placement level 2); the walks run from the CTA's global block."""
    csr, plan, cfg = workloads.grid_workload(240, 250, 4, 0.1, 4, seed=5, tree_kind=_abi.RC_TREE_WILSON,
                                             cap_nodes=32000, cap_edges=65000)
    eng = _free_run_vs_oracle(csr, plan, cfg, 2)
    assert eng.info().spill == 2


def test_config4_wilson_200k_nodes():
    """BASELINE configs[3] (state-house scale): 200k-node Delaunay graph, 50 districts, Wilson."""
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plan_delaunay200000_k50.npy")
    if not os.path.exists(path):
        pytest.skip("seed plan fixture not generated")
    csr, plan, cfg = workloads.config4(n_chains=8, seed=41)
    _free_run_vs_oracle(csr, plan, cfg, 3)


def test_edge_cases_status_words():
    """Inputs the reference rejects or chokes on map to status words, never to a hang:
    a plan whose district is disconnected (merged pair not connected), a merged pair larger
    than the engine's caps, a plan with a single district (no cut edges), out-of-range ids."""
    from gerrychain_b200.engine import Engine, EngineError

    csr, plan, cfg = workloads.config1(n_chains=4, seed=1)
    # (1) district 0 split into two far-apart blobs: every pair containing it is disconnected
    bad = plan.copy().reshape(20, 20)
    bad[:, :] = 1
    bad[0:2, 0:2] = 0
    bad[18:20, 18:20] = 0
    bad[10:20, 0:10] = 2
    bad[0:10, 10:20] = 3
    for tree_kind in (_abi.RC_TREE_MST, _abi.RC_TREE_WILSON):
        cfg2 = workloads.config1(n_chains=4, seed=1, tree_kind=tree_kind, cap_nodes=400, cap_edges=760)[2]
        cfg2.pop_lower, cfg2.pop_upper = 1.0, 0.0  # no bounds: the plan is unbalanced on purpose
        cfg2.epsilon = 0.9
        eng = Engine(csr, cfg2, bad.reshape(-1))
        eng.run(30).synchronize()
        st = eng.status.cpu().numpy()
        assert set(st.tolist()) <= {_abi.RC_OK, _abi.RC_ERR_DISCONNECTED}
        assert (st == _abi.RC_ERR_DISCONNECTED).any()
    # (2) caps smaller than a merged pair
    cfg3 = workloads.config1(n_chains=2, seed=1, cap_nodes=64, cap_edges=128)[2]
    eng = Engine(csr, cfg3, plan)
    eng.run(1).synchronize()
    assert (eng.status.cpu().numpy() == _abi.RC_ERR_CAPACITY).all()
    with pytest.raises(EngineError):
        eng.check_status()
    # (3) one district only: no cut edges to pick from (random.choice(()) in the reference)
    cfg4 = workloads.config1(n_chains=2, seed=1)[2]
    eng = Engine(csr, cfg4, np.zeros(csr.n_nodes, dtype=np.uint8))
    eng.run(1).synchronize()
    assert (eng.status.cpu().numpy() == _abi.RC_ERR_NO_CUT_EDGES).all()
    # (4) ids outside [0, k)
    with pytest.raises(ValueError):
        Engine(csr, cfg4, np.full(csr.n_nodes, 9, dtype=np.uint8))
    # (5) single chain, many districts (k = 100 on a 20x20 grid: 4-node districts, pairs of 8 nodes)
    k = 100
    tiny = (np.arange(400).reshape(20, 20) // 40 * 10 + np.arange(400).reshape(20, 20) % 20 // 2).astype(np.uint8)
    from gerrychain_b200.config import EngineConfig
    cfg5 = EngineConfig(n_districts=k, n_chains=3, pop_target=4.0, epsilon=0.3, seed=3)
    _free_run_vs_oracle(csr, tiny.reshape(-1), cfg5, 40)


@pytest.mark.parametrize("name", _golden.seedplan_names())
def test_seed_plan_replay_bit_exact_with_reference(name):
    """recursive_tree_part (tree.py:971-1085): the reference's recorded weights, roots and cut
    choices replayed on the device give the reference's seed plan."""
    from gerrychain_b200.engine import Engine

    csr, cfg, rec, plan, _ = _golden.load_seedplan(name)
    cfg.cap_nodes, cfg.cap_edges = csr.n_nodes, csr.n_edges
    eng = Engine(csr, cfg, np.zeros(csr.n_nodes, dtype=np.uint8))
    eng.seed_replay(rec)
    assert int(eng.status.cpu().numpy()[0]) == 0, _abi.STATUS_NAMES.get(int(eng.status.cpu().numpy()[0]))
    assert np.array_equal(eng.assignments()[0], plan)
    _naive_check(csr, cfg.n_districts, eng.assignments(), eng.tally.cpu().numpy(),
                 eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())


def test_seed_plans_free_running_bit_exact_with_oracle():
    from oracle import oracle
    from gerrychain_b200.engine import Engine

    for name, n_plans in (("seedplan_grid20_k8", 24), ("seedplan_delaunay600_k5", 16)):
        csr, cfg, _, _, _ = _golden.load_seedplan(name)
        cfg.n_chains, cfg.seed = n_plans, 5150
        cfg.cap_nodes, cfg.cap_edges = csr.n_nodes, csr.n_edges
        eng = Engine(csr, cfg, np.zeros(csr.n_nodes, dtype=np.uint8))
        eng.seed(10000).synchronize()
        status, plans, tally, counters = oracle.seed_plans(csr, cfg, n_plans=n_plans)
        assert np.array_equal(eng.status.cpu().numpy(), status) and (status == 0).all()
        assert np.array_equal(eng.assignments(), plans)
        assert np.array_equal(eng.tally.cpu().numpy(), tally)
        assert np.array_equal(eng.counters.cpu().numpy()[:, _abi.RC_CTR_TREES], counters[:, _abi.RC_CTR_TREES])
        _naive_check(csr, cfg.n_districts, eng.assignments(), eng.tally.cpu().numpy(),
                     eng.cut_mask.cpu().numpy(), eng.cut_count.cpu().numpy())
        # a seeded engine is ready to run chains from its own plans
        cfg2 = cfg
        eng.run(5).synchronize()
        assert (eng.status.cpu().numpy() == 0).all()


def test_seed_plans_config3_scale_spill_mode():
    """9k-node Delaunay graph, 18 districts: every split works on the whole remainder, so
    the seed kernel runs in spill mode; plans must be balanced and contiguous."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from gerrychain_b200.engine import Engine

    csr, _, cfg = workloads.config3(n_chains=8, seed=8)
    cfg.cap_nodes, cfg.cap_edges = csr.n_nodes, csr.n_edges
    eng = Engine(csr, cfg, np.zeros(csr.n_nodes, dtype=np.uint8))
    eng.seed(10000).synchronize()
    assert (eng.status.cpu().numpy() == 0).all()
    tally = eng.tally.cpu().numpy()
    assert (tally >= cfg.pop_lower).all() and (tally <= cfg.pop_upper).all()
    uv = csr.edge_uv
    for pl in eng.assignments():
        same = pl[uv[:, 0]] == pl[uv[:, 1]]
        m = coo_matrix((np.ones(same.sum()), (uv[same, 0], uv[same, 1])), shape=(csr.n_nodes,) * 2)
        assert connected_components(m, directed=False)[0] == cfg.n_districts


def test_histograms():
    csr, plan, cfg = workloads.config1(n_chains=128, seed=2)
    eng = _engine(csr, cfg, plan)
    eng.run(20)
    cut, dev = eng.histograms(256, 64, 0.001)
    eng.synchronize()
    cc = eng.cut_count.cpu().numpy()
    assert np.array_equal(cut.cpu().numpy(), np.bincount(np.minimum(cc, 255), minlength=256))
    assert int(dev.sum()) == 128
