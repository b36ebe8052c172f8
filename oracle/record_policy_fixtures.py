#!/usr/bin/env python
"""Record reference ReCom steps on the two branches of the cut choice that the replay_* fixtures
do not reach (tests/golden/policy_*.npz; CPU oracle tests only, so the names stay out of the
replay_* glob the device tests iterate over).

TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference):

    python oracle/record_policy_fixtures.py

* more than three region columns: _region_preferred_max_weight_choice falls back to
  _max_weight_choice over ALL balanced cuts (/root/reference/gerrychain/tree.py:548-549);
* ``region_surcharge={}``: no surcharge on any weight, yet the cut is the one of maximum weight,
  not a uniform draw (tree.py:548 ``len(region_surcharge) == 0``)."""

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

from record_reference import REFERENCE, Recorder  # noqa: E402


def main():
    import networkx

    from gerrychain_b200 import synthetic as syn

    gold = os.path.join(REPO, "tests", "golden")
    with open(os.path.join(REFERENCE, "tests/graphs_for_test/8x8_with_muni.json")) as f:
        data = json.load(f)
    g8 = networkx.readwrite.json_graph.adjacency_graph(data)
    for v in g8:
        g8.nodes[v]["row"] = int(g8.nodes[v]["district"])  # a fourth region column: the seed plan's rows
    plan8 = {v: int(g8.nodes[v]["district"]) - 1 for v in g8}
    Recorder(g8, plan8, "TOTPOP", 8, 0.25,
             region_surcharge={"muni": 0.3, "county": 0.2, "county2": 0.7, "row": 0.1}).record(
        80, os.path.join(gold, "policy_8x8_region4.npz"), seed=12)
    gd = syn.grid_graph(10, 10)
    Recorder(gd, syn.grid_quadrants(10, 10), "population", 4, 0.1, region_surcharge={}).record(
        60, os.path.join(gold, "policy_grid10_empty_surcharge.npz"), seed=13)


if __name__ == "__main__":
    main()
