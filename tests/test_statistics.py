"""Per-plan statistics: region splits and contiguity.

The fixtures tests/golden/stats_*.npz hold what the unmodified reference reports for 120
plans per graph (oracle/record_statistics.py): ``total_reg_splits`` and
``compute_county_splits`` (updaters/county_splits.py), ``contiguous`` and
``contiguous_components`` (constraints/contiguity.py).  CPU: the oracle against them.
GPU: ``rc_engine_region_splits`` / ``rc_engine_contiguity`` against them and the oracle, all
plans resident as the chains of one engine.
"""

import glob
import json
import os

import numpy as np
import pytest

import _golden
from gerrychain_b200 import workloads
from gerrychain_b200.config import EngineConfig
from gerrychain_b200.csr import csr_from_edges

CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(_golden.GOLDEN, "stats_*.npz")))


def _load(name):
    z = np.load(os.path.join(_golden.GOLDEN, f"stats_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    csr = csr_from_edges(int(z["pop"].shape[0]), z["edge_uv"], z["pop"])
    return z, meta, csr


def test_fixtures_present():
    assert CASES == ["8x8_muni", "delaunay600"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_statistics_match_reference(name):
    from oracle import oracle

    z, meta, csr = _load(name)
    k = meta["k"]
    for i, plan in enumerate(z["plans"]):
        for c, R in enumerate(meta["n_regions"]):
            got = oracle.region_splits(csr, k, plan, z["region_ids"][c], R)
            assert got == (z["edge_splits"][i, c], z["split_regions"][i, c], z["pieces"][i, c]), (i, c)
        pieces = oracle.contiguity(csr, k, plan)
        assert np.array_equal(pieces, z["district_pieces"][i])
        assert bool((pieces <= 1).all()) == bool(z["is_contiguous"][i])
    assert 0 < z["is_contiguous"].sum() < len(z["plans"])  # both outcomes are exercised


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_statistics_match_reference(name):
    from gerrychain_b200.engine import Engine

    z, meta, csr = _load(name)
    plans = z["plans"]
    k = meta["k"]
    cfg = EngineConfig(n_districts=k, n_chains=len(plans), pop_target=float(csr.pop.sum()) / k, epsilon=0.5)
    eng = Engine(csr, cfg, plans)
    for c, R in enumerate(meta["n_regions"]):
        got = eng.region_splits(z["region_ids"][c], R).cpu().numpy()
        assert np.array_equal(got[:, 0], z["edge_splits"][:, c])
        assert np.array_equal(got[:, 1], z["split_regions"][:, c])
        assert np.array_equal(got[:, 2], z["pieces"][:, c])
    cc = eng.contiguity().cpu().numpy()
    assert np.array_equal(cc, z["district_pieces"])
    assert np.array_equal((cc <= 1).all(axis=1), z["is_contiguous"].astype(bool))
    # Election tallies (updaters/election.py): per-district vote totals of both parties
    for party in range(2):
        got = eng.tally_column(z["votes"][party]).cpu().numpy()
        assert np.array_equal(got, z["vote_counts"][:, party, :])


@pytest.mark.gpu
@pytest.mark.parametrize("workload,chains,steps", [("config2", 296, 40), ("config5", 296, 40)])
def test_recom_keeps_plans_contiguous_and_statistics_match_oracle(workload, chains, steps):
    """Full-size property: every state a ReCom chain yields is contiguous
    (constraints/contiguity.py:169-181 holds by construction), and the device statistics of
    the evolved plans equal the oracle's on the same plans."""
    from gerrychain_b200.engine import Engine
    from oracle import oracle

    csr, plan, cfg = getattr(workloads, workload)(n_chains=chains, seed=3)
    eng = Engine(csr, cfg, plan)
    eng.run(steps)
    eng.check_status()
    cc = eng.contiguity().cpu().numpy()
    assert (cc == 1).all()
    assign = eng.assignments()
    rng = np.random.default_rng(0)
    region = (csr.region_ids[0] if csr.n_region_cols else rng.integers(0, 40, csr.n_nodes)).astype(np.int32)
    R = int(region.max()) + 1
    got = eng.region_splits(region, R).cpu().numpy()
    for ch in range(0, chains, 37):
        assert tuple(got[ch]) == oracle.region_splits(csr, cfg.n_districts, assign[ch], region, R)
        assert np.array_equal(cc[ch], oracle.contiguity(csr, cfg.n_districts, assign[ch]))


@pytest.mark.gpu
def test_contiguity_global_scratch_path_200k_nodes():
    """N = 200 000 nodes: the union-find parents do not fit in shared memory."""
    from gerrychain_b200.engine import Engine
    from oracle import oracle

    csr, plan, cfg = workloads.config4(n_chains=3, seed=1)
    plans = np.repeat(plan[None, :], 3, axis=0).copy()
    rng = np.random.default_rng(2)
    idx = rng.integers(0, csr.n_nodes, 50)
    plans[1, idx] = rng.integers(0, cfg.n_districts, 50)  # chain 1: scattered single nodes
    eng = Engine(csr, cfg, plans)
    cc = eng.contiguity().cpu().numpy()
    assert (cc[0] == 1).all()
    for ch in range(3):
        assert np.array_equal(cc[ch], oracle.contiguity(csr, cfg.n_districts, plans[ch]))
    assert cc[1].sum() > cfg.n_districts
