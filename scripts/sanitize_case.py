"""Tiny workload for compute-sanitizer: a few chains, MST + Wilson + seed plans + spill."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gerrychain_b200 import workloads, _abi
from gerrychain_b200.engine import Engine

which = sys.argv[1] if len(sys.argv) > 1 else "mst"
if which == "mst":
    csr, plan, cfg = workloads.config1(n_chains=6, seed=3)
    e = Engine(csr, cfg, plan); e.run(4).synchronize(); e.check_status()
elif which == "wilson":
    csr, plan, cfg = workloads.config1(n_chains=4, seed=3, tree_kind=_abi.RC_TREE_WILSON)
    e = Engine(csr, cfg, plan); e.run(3).synchronize(); e.check_status()
elif which == "seed":
    csr, plan, cfg = workloads.config1(n_chains=3, seed=3, cap_nodes=400, cap_edges=760)
    e = Engine(csr, cfg, plan); e.seed().synchronize(); e.check_status(); e.run(2).synchronize()
elif which == "spill":
    csr, plan, cfg = workloads.config3(n_chains=3, seed=3, cap_nodes=9000, cap_edges=26976)
    e = Engine(csr, cfg, plan); e.run(2).synchronize(); e.check_status()
print("ok", which)
