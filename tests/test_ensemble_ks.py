"""Free-running ensembles against the reference (north star): two-sample KS, alpha = 0.01,
on the number of cut edges and on the population deviation, sampled ACROSS independent
chains at a fixed step.  The reference ensembles are tests/golden/ensemble_*.npz, written by
oracle/make_ensemble_fixtures.py from the unmodified reference (384 chains each).

CPU: the oracle port's free-running mode vs the reference (pins the counter-based RNG
schedule statistically).  GPU: the CUDA engine vs the reference."""

import json
import os

import numpy as np
import pytest
from scipy.stats import ks_2samp

from gerrychain_b200 import _abi, synthetic
from gerrychain_b200.config import EngineConfig, cut_policy_for
from gerrychain_b200.csr import csr_from_networkx

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALPHA = 0.01
CASES = ["grid20_mst", "grid20_wilson", "grid12_region", "grid20_cutaccept", "grid12_reversible", "delaunay600_mst"]
# BASELINE configs 3 and 5 at their stated size (9k nodes, 18 districts, 67 counties): the
# reference ensembles took ~10 core-minutes each; the oracle leg runs fewer chains to stay
# within the CPU suite's budget, the engine leg runs 4096.
CASES_9K = ["delaunay9000_region", "delaunay9000_mst", "delaunay9000_wilson"]


def _case(name, n_chains, seed):
    z = np.load(os.path.join(GOLDEN, f"ensemble_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    if meta.get("graph") == "delaunay":
        g = synthetic.delaunay_graph(meta["n"], seed=0)
        m = n = 0
    else:
        m, n = meta["dims"]
        g = synthetic.grid_graph(m, n)
    regions = None
    if meta.get("counties"):
        synthetic.voronoi_counties(g, meta["counties"], seed=1, key="county")
        regions = {"county": 0.5}
    elif meta.get("region"):
        for (i, j) in g.nodes:
            g.nodes[(i, j)]["county"] = (i // (m // 3)) * 3 + j // (n // 3)
        regions = {"county": 0.5}
    csr = csr_from_networkx(g, "population", regions)
    k = int(z["plan"].max()) + 1
    ideal = float(csr.pop.sum()) / k
    eps = meta["eps"]
    tree_kind = _abi.RC_TREE_WILSON if meta["tree"] == "wilson" else _abi.RC_TREE_MST
    cfg = EngineConfig(n_districts=k, n_chains=n_chains, pop_target=ideal, epsilon=eps,
                       pop_lower=(1 - eps) * ideal, pop_upper=(1 + eps) * ideal, seed=seed,
                       tree_kind=tree_kind, cut_policy=cut_policy_for(regions, tree_kind),
                       accept_rule=_abi.RC_ACCEPT_CUT_EDGE if meta.get("accept") == "cut_edge" else _abi.RC_ACCEPT_ALWAYS,
                       proposal_kind=_abi.RC_PROPOSAL_REVERSIBLE if meta.get("reversible_M") else _abi.RC_PROPOSAL_RECOM,
                       rev_M=meta.get("reversible_M", 1))
    return z, meta, g, csr, cfg, ideal


def _county_splits(g, csr, assign, k):
    """sum over districts of (distinct counties touched - 1), per chain."""
    county = np.array([g.nodes[v]["county"] for v in csr.node_labels], dtype=np.int64)
    nc = int(county.max()) + 1
    out = np.zeros(assign.shape[0], dtype=np.int64)
    for c in range(assign.shape[0]):
        pairs = np.unique(assign[c].astype(np.int64) * nc + county)
        out[c] = len(pairs) - len(np.unique(pairs // nc))
    return out


def _compare(z, meta, g, csr, cfg, ideal, cut_count, tally, assign):
    ref_cut = z["cut_edges"]
    ref_dev = np.abs(z["population"] - ideal).max(axis=1) / ideal
    our_dev = np.abs(tally - ideal).max(axis=1) / ideal
    p_cut = ks_2samp(ref_cut, cut_count).pvalue
    p_dev = ks_2samp(ref_dev, our_dev).pvalue
    p_pool = ks_2samp((z["population"].ravel() - ideal) / ideal, (tally.ravel() - ideal) / ideal).pvalue
    assert p_cut > ALPHA, ("cut edges", p_cut, ref_cut.mean(), cut_count.mean())
    assert p_dev > ALPHA, ("max |deviation|", p_dev)
    assert p_pool > ALPHA, ("pooled deviation", p_pool)
    if meta.get("region"):
        ours = _county_splits(g, csr, assign, cfg.n_districts)
        p_split = ks_2samp(z["county_splits"], ours).pvalue
        assert p_split > ALPHA, ("county splits", p_split, z["county_splits"].mean(), ours.mean())


@pytest.mark.parametrize("name", CASES)
def test_oracle_ensemble_matches_reference(name):
    from oracle import oracle

    z, meta, g, csr, cfg, ideal = _case(name, 1536, seed=11)
    ora = oracle.OracleChains(csr, cfg, z["plan"]).run(meta["steps"], n_threads=os.cpu_count() or 1)
    assert (ora.status == 0).all()
    _compare(z, meta, g, csr, cfg, ideal, ora.cut_count, ora.tally, ora.assignments())


@pytest.mark.parametrize("name", CASES_9K)
def test_oracle_ensemble_matches_reference_9k(name):
    from oracle import oracle

    z, meta, g, csr, cfg, ideal = _case(name, 512 if meta_tree(name) == "mst" else 256, seed=11)
    ora = oracle.OracleChains(csr, cfg, z["plan"]).run(meta["steps"], n_threads=os.cpu_count() or 1)
    assert (ora.status == 0).all()
    _compare(z, meta, g, csr, cfg, ideal, ora.cut_count, ora.tally, ora.assignments())


def test_oracle_ensemble_matches_reference_200k():
    """BASELINE config 4 at its stated size (Delaunay 200k nodes, 50 districts, eps 0.05,
    uniform_spanning_tree): 128 chains x 20 steps of the unmodified reference
    (oracle/make_ensemble_fixtures.py delaunay200000_wilson, ~1.7 s per reference step) against the
    oracle from the same committed seed plan."""
    from oracle import oracle
    from gerrychain_b200 import workloads

    z = np.load(os.path.join(GOLDEN, "ensemble_delaunay200000_wilson.npz"))
    meta = json.loads(str(z["meta"]))
    csr, plan, cfg = workloads.config4(n_chains=256, seed=11)
    assert csr.n_nodes == meta["n"] and cfg.n_districts == meta["k"] and cfg.epsilon == meta["eps"]
    assert np.array_equal(plan, z["plan"])
    ora = oracle.OracleChains(csr, cfg, plan).run(meta["steps"], n_threads=os.cpu_count() or 1)
    assert (ora.status == 0).all()
    _compare(z, meta, None, csr, cfg, cfg.pop_target, ora.cut_count, ora.tally, ora.assignments())


def meta_tree(name):
    return "wilson" if "wilson" in name else "mst"


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES + CASES_9K)
def test_engine_ensemble_matches_reference(name):
    from gerrychain_b200.engine import Engine

    z, meta, g, csr, cfg, ideal = _case(name, 4096 if "wilson" not in name or "9000" not in name else 2048, seed=12)
    eng = Engine(csr, cfg, z["plan"])
    eng.run(meta["steps"]).synchronize()
    eng.check_status()
    _compare(z, meta, g, csr, cfg, ideal, eng.cut_count.cpu().numpy(), eng.tally.cpu().numpy(),
             eng.assignments())
