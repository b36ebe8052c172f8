"""CPU tests: the C oracle against the reference's own known answers and recorded steps.

Known-answer trees: /root/reference/tests/test_tree.py:354-385.
Recorded steps: tests/golden/replay_*.npz (oracle/record_reference.py, unmodified reference).
"""

import numpy as np
import pytest

import _golden
from gerrychain_b200 import _abi
from oracle import oracle


def test_philox_known_answers(oracle_lib):
    # Random123 kat_vectors, philox4x32-10
    assert list(oracle.philox(0, 0, [0, 0, 0, 0])) == [
        0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(oracle.philox(0xFFFFFFFF, 0xFFFFFFFF, [0xFFFFFFFF] * 4)) == [
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(oracle.philox(0xA4093822, 0x299F31D0,
                              [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344])) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_find_balanced_cuts_known_answer(oracle_lib):
    # reference tests/test_tree.py:354-372: 9-node tree, unit pops, ideal 4.5, eps 0.5
    tree = [(0, 1), (1, 2), (1, 4), (3, 4), (4, 5), (3, 6), (6, 7), (6, 8)]
    for root in (1, 3, 4, 6):
        cuts = oracle.balanced_cuts_of_tree(9, tree, [1] * 9, 4.5, 0.5, root)
        assert {tuple(sorted(c)) for c in cuts} == {(1, 4), (3, 4), (3, 6)}


def test_no_balanced_cuts_when_one_side_okay(oracle_lib):
    # reference tests/test_tree.py:375-385: path, pops {4,4,3,3,3}, ideal 10, eps 0.1
    tree = [(0, 1), (1, 2), (2, 3), (3, 4)]
    for root in (1, 2, 3):
        assert oracle.balanced_cuts_of_tree(5, tree, [4, 4, 3, 3, 3], 10, 0.1, root) == []


def test_mst_matches_scipy(oracle_lib):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import minimum_spanning_tree

    g = _golden.load("replay_delaunay2400_mst")
    rng = np.random.default_rng(5)
    for _ in range(3):
        w = rng.random(g.csr.n_edges) + 0.01
        uv = g.csr.edge_uv
        m = coo_matrix((w, (uv[:, 0], uv[:, 1])), shape=(g.csr.n_nodes,) * 2).tocsr()
        t = minimum_spanning_tree(m).tocoo()
        look = g.csr.edge_id_lookup()
        want = sorted(look[(min(a, b), max(a, b))] for a, b in zip(t.row, t.col))
        got = oracle.mst_of_graph(g.csr, w)
        assert list(got) == want


@pytest.mark.parametrize("name", _golden.names())
def test_replay_matches_reference(oracle_lib, name):
    g = _golden.load(name)
    ch = oracle.OracleChains(g.csr, g.cfg, g.init_assignment)
    rc, trace, info = ch.replay(g.rec)
    assert rc == 0, _abi.STATUS_NAMES.get(rc)
    assert np.array_equal(trace, g.assignment), "assignment differs from the reference"
    assert np.array_equal(info[:, 2], g.n_cut_edges)
    # the balanced-cut set the reference saw in its successful attempt
    assert np.array_equal(info[:, 1], np.diff(g.balanced_ptr))
    # trees used == trees the reference drew (earlier trees had no balanced cut)
    assert np.array_equal(info[:, 0], np.diff(g.tree_ptr))
    assert np.array_equal(ch.tally[0], g.population[-1])
    # incremental cut mask == naive recomputation (updaters/test_cut_edges.py:144-161)
    uv = g.csr.edge_uv
    a = trace[-1]
    naive = np.nonzero(a[uv[:, 0]] != a[uv[:, 1]])[0]
    bits = np.unpackbits(ch.cut_mask[0].view(np.uint8), bitorder="little")[: g.csr.n_edges]
    assert np.array_equal(np.nonzero(bits)[0], naive)


@pytest.mark.parametrize("name", [n for n in _golden.names() if "region" in n])
def test_region_cut_policy_matches_reference(oracle_lib, name):
    """cut_edge not injected: the oracle's own power-set / max-weight choice must pick the
    reference's cut (tree.py:501-578)."""
    g = _golden.load(name, inject_cut=False)
    ch = oracle.OracleChains(g.csr, g.cfg, g.init_assignment)
    rc, trace, info = ch.replay(g.rec)
    assert rc == 0
    assert np.array_equal(trace, g.assignment)


@pytest.mark.parametrize("name", _golden.seedplan_names())
def test_oracle_seed_plan_replays_reference_recursive_tree_part(name):
    """tree.py:971-1085: fed the reference's recorded weights / roots / cut choices, the
    oracle's recursive_tree_part returns the reference's plan, split by split."""
    csr, cfg, rec, plan, _ = _golden.load_seedplan(name)
    status, plans, tally, _ = oracle.seed_plans(csr, cfg, rec=rec)
    assert status[0] == 0, _abi.STATUS_NAMES.get(int(status[0]))
    assert np.array_equal(plans[0], plan)
    assert np.array_equal(tally[0], np.bincount(plan, weights=csr.pop, minlength=cfg.n_districts).astype(np.int64))


def test_oracle_seed_plans_free_running_are_valid():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    csr, cfg, _, _, _ = _golden.load_seedplan("seedplan_grid20_k8")
    cfg.seed = 77
    status, plans, tally, _ = oracle.seed_plans(csr, cfg, n_plans=20)
    assert (status == 0).all()
    lo, hi = cfg.pop_target * (1 - cfg.epsilon), cfg.pop_target * (1 + cfg.epsilon)
    assert (tally >= lo).all() and (tally <= hi).all()
    uv = csr.edge_uv
    for pl in plans:
        assert set(pl.tolist()) == set(range(cfg.n_districts))
        same = pl[uv[:, 0]] == pl[uv[:, 1]]
        m = coo_matrix((np.ones(same.sum()), (uv[same, 0], uv[same, 1])), shape=(csr.n_nodes,) * 2)
        assert connected_components(m, directed=False)[0] == cfg.n_districts
    assert len({pl.tobytes() for pl in plans}) == 20


@pytest.mark.parametrize("name", ["reversible_grid8", "reversible_grid12"])
def test_reversible_replay_matches_reference(name):
    """reversible_recom (tree_proposals.py:149-267) pinned bit-exactly, not only statistically: the
    reference's recorded calls (oracle/record_reversible.py: ordered pair, edge, walk stacks, BFS
    root, chosen cut, acceptance draw) replayed through the oracle give the reference's plan after
    every call, the same number of balanced cuts per tree and the same outcome (self-loop kinds /
    move).  The device runs the same code path bit-exactly with the oracle on Philox draws
    (tests/test_gpu_parity.py::test_reversible_recom_bit_exact_with_oracle)."""
    import json
    import os

    from oracle import oracle
    from gerrychain_b200.config import EngineConfig
    from gerrychain_b200.csr import csr_from_edges

    z = np.load(os.path.join(_golden.GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    N = int(z["pop"].shape[0])
    csr = csr_from_edges(N, z["edge_uv"], z["pop"])
    assert np.array_equal(csr.edge_uv, z["edge_uv"])
    cfg = EngineConfig(n_districts=meta["k"], n_chains=1, pop_target=meta["pop_target"], epsilon=meta["epsilon"],
                       tree_kind=_abi.RC_TREE_WILSON, proposal_kind=_abi.RC_PROPOSAL_REVERSIBLE, rev_M=meta["M"])
    S = meta["n_steps"]
    steps = np.zeros(S, dtype=oracle.REV_STEP_DTYPE)
    for f in ("d_out", "d_in", "pair_edge", "root", "cut_edge", "tree_idx", "accept_u"):
        steps[f] = z[f]
    T = len(z["wilson_root"])
    sp = np.zeros((T, N + 1), dtype=np.int64)  # dense per-tree stack pointers, as _golden.load builds them
    wn_ptr, wn_node, ws_ptr = z["wn_ptr"], z["wn_node"], z["ws_ptr"]
    for t in range(T):
        lo, hi = int(wn_ptr[t]), int(wn_ptr[t + 1])
        cnt = np.zeros(N, dtype=np.int64)
        cnt[wn_node[lo:hi]] = ws_ptr[lo + 1:hi + 1] - ws_ptr[lo:hi]
        sp[t, 0] = ws_ptr[lo]
        sp[t, 1:] = ws_ptr[lo] + np.cumsum(cnt)
    rc, trace, info, tally, cut_count = oracle.replay_reversible(csr, cfg, z["init_assignment"], steps,
                                                                z["wilson_root"], sp, z["ws_val"])
    assert rc == 0, _abi.STATUS_NAMES.get(rc, rc)
    assert np.array_equal(trace, z["assignment"])
    assert np.array_equal(info[:, 3], z["outcome"])
    had_tree = z["outcome"] >= oracle.REV_NO_CUT
    assert np.array_equal(info[had_tree, 1], z["n_balanced"][had_tree])
    assert np.array_equal(info[:, 0], had_tree.astype(np.int32))
    assert cut_count == int(z["n_cut_edges"][-1])
    assert (z["outcome"] == oracle.REV_MOVED).sum() >= 5  # the record does contain accepted moves
    last = z["assignment"][-1].astype(np.int64)
    assert np.array_equal(tally, np.bincount(last, weights=csr.pop, minlength=meta["k"]).astype(np.int64))


@pytest.mark.parametrize("name", ["policy_8x8_region4", "policy_grid10_empty_surcharge"])
def test_max_weight_cut_policy_matches_reference(oracle_lib, name):
    """The branches of the cut choice the replay_* fixtures do not reach (tree.py:548-549): more
    than three region columns, and ``region_surcharge={}`` - in both the reference takes the
    balanced cut of maximum weight (oracle/record_policy_fixtures.py).  Replayed with and without
    the reference's chosen cut injected."""
    for inject in (True, False):
        g = _golden.load(name, inject_cut=inject)
        if "empty" in name:  # the fixture stores no region column: {} -> _max_weight_choice
            assert g.cfg.cut_policy == _abi.RC_CUT_UNIFORM
            g.cfg.cut_policy = _abi.RC_CUT_MAX_WEIGHT
        else:
            assert g.csr.n_region_cols == 4
        assert g.cfg.cut_policy == _abi.RC_CUT_MAX_WEIGHT
        ch = oracle.OracleChains(g.csr, g.cfg, g.init_assignment)
        rc, trace, info = ch.replay(g.rec)
        assert rc == 0, _abi.STATUS_NAMES.get(rc)
        assert np.array_equal(trace, g.assignment)
        assert np.array_equal(info[:, 1], np.diff(g.balanced_ptr))
    # the choice is not vacuous: most steps offered more than one balanced cut
    assert (np.diff(g.balanced_ptr) > 1).mean() > 0.3
