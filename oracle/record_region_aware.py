#!/usr/bin/env python
"""Run the reference's OWN region-aware ensemble test chains and store what they report
(tests/golden/region_aware_reference.json).

TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference):

    python oracle/record_region_aware.py        # ~8 minutes on 8 cores

The chains are /root/reference/tests/test_region_aware.py: run_chain_single, called exactly as
test_region_aware_muni (:82-99), test_region_aware_county (:137-155) and
test_region_aware_muni_reselect (:115-134) call it - one process per chain, random.seed(chain
number) - only with more chains than the 30 / 100 the reference draws, so that the histograms of
"regions split at the last step" are sharp enough for a two-sample test
(tests/test_region_aware_mirror.py)."""

import json
import os
import sys
import warnings
from concurrent.futures import ProcessPoolExecutor
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("GERRYCHAIN_REFERENCE", "/root/reference")
sys.path[:0] = [REFERENCE, os.path.join(HERE, "refshim"), os.path.join(REFERENCE, "tests")]

CASES = {
    "county2": dict(category="county2", surcharge=0.8, steps=5000, max_attempts=100000, reselect=False, n=500, n_regions=8),
    "muni": dict(category="muni", surcharge=0.5, steps=5000, max_attempts=100000, reselect=False, n=240, n_regions=16),
    "muni_reselect": dict(category="muni", surcharge=1.0, steps=500, max_attempts=100, reselect=True, n=480, n_regions=16),
}


def main():
    os.chdir(REFERENCE)  # run_chain_single opens tests/graphs_for_test/8x8_with_muni.json relative to the repo root
    warnings.filterwarnings("ignore")
    from test_region_aware import run_chain_single

    out = {"generator": "oracle/record_region_aware.py (the reference's own tests/test_region_aware.py: "
                        "run_chain_single, one process per chain, random.seed(chain))",
           "graph": "tests/graphs_for_test/8x8_with_muni.json (carried by tests/golden/replay_8x8_region3.npz)",
           "cases": {}}
    with ProcessPoolExecutor(os.cpu_count()) as ex:
        for name, c in CASES.items():
            res = list(ex.map(partial(run_chain_single, category=c["category"], steps=c["steps"],
                                      surcharge=c["surcharge"], max_attempts=c["max_attempts"],
                                      reselect=c["reselect"]), range(c["n"])))
            out["cases"][name] = dict(category=c["category"], surcharge=c["surcharge"], steps=c["steps"],
                                      max_attempts=c["max_attempts"], reselect=c["reselect"],
                                      seeds=f"0..{c['n'] - 1}", n_regions=c["n_regions"],
                                      splits_histogram=np.bincount(res).tolist())
            print(name, out["cases"][name]["splits_histogram"], flush=True)
    with open(os.path.join(os.path.dirname(HERE), "tests", "golden", "region_aware_reference.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
