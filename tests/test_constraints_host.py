"""Scalar bound wrappers, validity helpers and Polsby-Popper norms of ``constraints``
(constraints/bounds.py:36-236, validity.py:125-186, constraints/compactness.py of the
reference): known answers, including the reference's quirks (the ``>=`` in ``UpperBound``'s
name, a falsy self-configured bound being configured again, the configuring call of
``WithinPercentRangeOfBounds`` answering True).  Host logic only."""

import math

import pytest

from gerrychain_b200 import constraints as C


def score(x):
    return x


class FakePartition(dict):
    def __init__(self, parent=None, parts=None, **updaters):
        super().__init__(**updaters)
        self.parent = parent
        self.parts = parts or {}

    @property
    def assignment(self):
        return self


def test_upper_and_lower_bound():
    up, low = C.UpperBound(score, 5), C.LowerBound(score, 5)
    assert [up(4), up(5), up(5.1)] == [True, True, False]
    assert [low(4.9), low(5), low(6)] == [False, True, True]
    assert up.__name__ == "UpperBound(score >= 5)" and repr(up) == "<UpperBound(score >= 5)>"
    assert low.__name__ == "LowerBound(score <= 5)"


def test_self_configuring_bounds():
    up = C.SelfConfiguringUpperBound(score)
    assert [up(3), up(4), up(2), up(3)] == [True, False, True, True] and up.bound == 3
    zero = C.SelfConfiguringUpperBound(score)
    assert zero(0) and zero(7) and zero.bound == 7  # bound 0 is falsy: configured again (bounds.py:139-141)
    low = C.SelfConfiguringLowerBound(score, epsilon=0.5)
    assert [low(3), low(2.5), low(2.49)] == [True, True, False] and low.bound == 2.5
    assert C.SelfConfiguringLowerBound(score).epsilon == 0.05
    assert low.__name__ == "SelfConfiguringLowerBound(score)"


def test_within_percent_range_of_bounds():
    w = C.WithinPercentRangeOfBounds(score, 10)
    assert w(100) is True  # the configuring call
    assert (w.lbound, w.ubound) == (90.0, 110.00000000000001)
    assert [w(109), w(111), w(90), w(89)] == [True, False, True, False]
    assert w.__name__ == "WithinPercentRangeOfBounds(score)"


def test_validator_accepts_the_wrappers_and_rejects_non_booleans():
    v = C.Validator([C.UpperBound(score, 5), C.LowerBound(score, 1)])
    assert v(3) is True and v(6) is False and v(0) is False
    with pytest.raises(TypeError):
        C.Validator([lambda p: 1])(3)


def test_districts_within_tolerance_and_vanishing_districts():
    p = FakePartition(population={0: 100, 1: 105, 2: 95})
    assert C.districts_within_tolerance(p, "population", 0.11) is True  # 10 <= 0.11 * 95
    assert C.districts_within_tolerance(p, "population", 0.1) is False
    assert C.districts_within_tolerance(p, "population", 11) is True   # >= 1 is read as percent
    assert C.districts_within_tolerance(p, "population", 10) is False
    assert C.no_vanishing_districts(FakePartition(parts={0: set()})) is True  # no parent: nothing moved
    child = FakePartition(parent=p, parts={0: {1, 2}, 1: set()})
    assert C.no_vanishing_districts(child) is False
    assert C.no_vanishing_districts(FakePartition(parent=p, parts={0: {1}, 1: {2}})) is True


def test_polsby_popper_norms():
    p = FakePartition(parts={0: {1}, 1: {2}, 2: {3}}, polsby_popper={0: 0.5, 1: 0.25, 2: 0.8})
    assert C.L1_reciprocal_polsby_popper(p) == 2 + 4 + 1.25
    assert math.isclose(C.L1_polsby_popper(p), 1.55, rel_tol=1e-15)
    assert math.isclose(C.L2_polsby_popper(p), math.sqrt(0.25 + 0.0625 + 0.64), rel_tol=1e-15)
    assert C.L_minus_1_polsby_popper(p) == 3 / 7.25
    assert isinstance(C.no_worse_L1_reciprocal_polsby_popper, C.SelfConfiguringUpperBound)
    assert isinstance(C.no_worse_L_minus_1_polsby_popper, C.SelfConfiguringLowerBound)
    assert isinstance(C.no_more_discontiguous, C.SelfConfiguringLowerBound)
    assert C.no_more_discontiguous.func is C.number_of_contiguous_parts


def test_partisan_scores_of_election_results():
    """metrics/partisan.py through ElectionResults (updaters/election.py:369-422): a five-district
    two-party election worked by hand - shares .60 .45 .30 .52 .48, mean .47, median .48."""
    import numpy as np

    from gerrychain_b200 import metrics
    from gerrychain_b200.updaters import Election, ElectionResults

    votes = np.array([[60, 45, 30, 52, 48], [40, 55, 70, 48, 52]])
    res = ElectionResults(Election("SEN", {"A": "a", "B": "b"}), votes, list(range(5)))
    assert math.isclose(res.mean_median(), 0.01, abs_tol=1e-15)
    assert math.isclose(res.mean_thirdian(), 0.01, abs_tol=1e-15)  # sorted shares[round(5 / 3)] = .48
    assert math.isclose(res.efficiency_gap(), -0.04, abs_tol=1e-15)  # (30 - 40 - 10 + 46 - 46) / 500
    assert math.isclose(res.partisan_bias(), 0.1, abs_tol=1e-15)    # 3 of 5 districts above the mean
    assert math.isclose(res.partisan_gini(), 0.032, abs_tol=1e-15)
    assert metrics.wasted_votes(60, 40) == (10, 40) and metrics.wasted_votes(50, 50) == (50, 0)
    assert res.wins("A") == res.seats("A") == 2 and res.votes("B") == res.counts("B") == (40, 55, 70, 48, 52)


def test_county_splits_and_refuse_new_splits():
    """updaters/county_splits.py:58-117 and validity.py:156-173 on a 4 x 4 lattice whose
    counties are the four quadrants: the start plan (left / right half) splits none of them; moving one
    cell makes a NEW_SPLIT, which stays NEW in later states (the reference never ages it),
    while a county split from the start is an OLD_SPLIT for good."""
    import networkx as nx

    from gerrychain_b200 import Partition
    from gerrychain_b200.updaters import CountyInfo, CountySplit, county_splits

    g = nx.grid_2d_graph(4, 4)
    for (i, j) in g.nodes:
        g.nodes[(i, j)]["county"] = "NSEW"[(i >= 2) + 2 * (j >= 2)]
    ups = {"splits": county_splits("splits", "county")}
    start = Partition(g, {v: int(v[0] >= 2) for v in g.nodes}, ups)
    info = start["splits"]
    assert set(info) == set("NSEW")
    assert all(isinstance(x, CountyInfo) and x.split == CountySplit.NOT_SPLIT for x in info.values())
    assert sorted(info["N"].nodes) == [(0, 0), (0, 1), (1, 0), (1, 1)] and info["N"].contains == {0}
    assert C.refuse_new_splits("splits")(start) is True

    child = start.flip({(1, 1): 1})
    assert child["splits"]["N"].split == CountySplit.NEW_SPLIT and child["splits"]["N"].contains == {0, 1}
    assert child["splits"]["S"].split == CountySplit.NOT_SPLIT
    assert C.refuse_new_splits("splits")(child) is False
    grandchild = child.flip({(0, 3): 1})
    assert grandchild["splits"]["N"].split == CountySplit.NEW_SPLIT  # parent was NEW, not OLD
    assert grandchild["splits"]["E"].split == CountySplit.NEW_SPLIT
    healed = grandchild.flip({(1, 1): 0})
    assert healed["splits"]["N"].split == CountySplit.NOT_SPLIT

    rows = Partition(g, {v: int(v[1] >= 1) for v in g.nodes}, ups)  # cuts N and S from the start
    assert rows["splits"]["N"].split == CountySplit.OLD_SPLIT and rows["splits"]["E"].split == CountySplit.NOT_SPLIT
    assert rows.flip({(3, 3): 0})["splits"]["N"].split == CountySplit.OLD_SPLIT
    assert C.refuse_new_splits("splits")(rows.flip({(0, 0): 1})) is True  # N stays an OLD split


def _three_by_three_with_counties():
    import networkx as nx

    g = nx.Graph()
    g.add_edges_from([(0, 1), (0, 3), (1, 2), (1, 4), (2, 5), (3, 4), (3, 6), (4, 5), (4, 7), (5, 8), (6, 7), (7, 8)])
    for v in g.nodes:
        g.nodes[v]["county"] = "abc"[v // 3]
        g.nodes[v]["pop"] = 1
    return g


ALL_SPLIT_SCORES = ["num_parts", "num_pieces", "naked_boundary", "shannon_entropy", "power_entropy",
                    "symmetric_entropy", "num_split_localities"]


def test_locality_splits_known_answers_of_the_reference():
    """The reference's own cases (tests/updaters/test_split_scores.py:84-110): rows of a 3 x 3
    lattice are the counties; districts = rows (nothing split) and districts = columns."""
    from gerrychain_b200 import Partition
    from gerrychain_b200.updaters import LocalitySplits

    g = _three_by_three_with_counties()
    ups = {"splits": LocalitySplits("splittings", "county", "pop", ALL_SPLIT_SCORES)}
    rows = Partition(g, {v: 1 + v // 3 for v in g.nodes}, ups)
    got = rows["splits"]
    assert got is rows.updaters["splits"].scores
    assert (got["num_parts"], got["num_pieces"], got["naked_boundary"], got["num_split_localities"]) == (3, 3, 0, 0)
    assert got["shannon_entropy"] == 0 and got["power_entropy"] == 0 and got["symmetric_entropy"] == 18
    assert rows.updaters["splits"].allowed_pieces == {"a": 1, "b": 1, "c": 1}

    ups = {"splits": LocalitySplits("splittings", "county", "pop", ALL_SPLIT_SCORES)}
    cols = Partition(g, {v: 1 + v % 3 for v in g.nodes}, ups)
    got = dict(cols["splits"])
    assert (got["num_parts"], got["num_pieces"], got["naked_boundary"], got["num_split_localities"]) == (9, 9, 6, 3)
    assert math.isclose(got["shannon_entropy"], math.log(3), rel_tol=1e-12)          # 1 < . < 1.2 there
    assert math.isclose(got["power_entropy"], 3 * 3 * (3 * (1 / 3) ** 0.95 - 1), rel_tol=1e-12)  # .5 < . < .6
    assert math.isclose(got["symmetric_entropy"], 2 * 3 * 3 * 3 * math.sqrt(1 / 3), rel_tol=1e-12)  # 31 < . < 32
    assert LocalitySplits("s", "county", "pop").scores == {"num_parts": None}
