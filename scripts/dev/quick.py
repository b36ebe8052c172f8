#!/usr/bin/env python
"""Dev tool: quick resident-state throughput of one workload variant.
   python scripts/dev/quick.py config3 --wilson --walkers 4 --threads 192 --chains 1024 --S 64 --reps 3"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
from gerrychain_b200 import _abi, workloads
from gerrychain_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--wilson", action="store_true")
ap.add_argument("--walkers", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--S", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cap-nodes", type=int, default=0)
ap.add_argument("--cap-edges", type=int, default=0)
a = ap.parse_args()
kw = dict(n_chains=a.chains, seed=2024)
if a.wilson: kw["tree_kind"] = _abi.RC_TREE_WILSON
if a.walkers: kw["wilson_walkers"] = a.walkers
if a.threads: kw["threads_per_cta"] = a.threads
if a.cap_nodes: kw["cap_nodes"] = a.cap_nodes
if a.cap_edges: kw["cap_edges"] = a.cap_edges
csr, plan, cfg = getattr(workloads, a.workload)(**kw)
eng = Engine(csr, cfg, plan)
info = eng.info()
eng.run(4).synchronize(); eng.check_status()
best = 0
tot_ms = 0.0
all_ms = []
c0 = eng.counters.sum(0).cpu().numpy()
for _ in range(a.reps):
    eng.run(a.S); ms = eng.last_run_ms()
    best = max(best, a.chains * a.S / ms * 1e3)
    tot_ms += ms
    all_ms.append(round(ms, 1))
eng.check_status()
c = eng.counters.sum(0).cpu().numpy() - c0
steps = a.chains * a.S * a.reps
print(f"{a.workload}{' wilson' if a.wilson else ''} W={info.wilson_walkers} thr={info.threads_per_cta} ctas/sm={info.ctas_per_sm} "
      f"smem={info.smem_bytes} spill={info.spill} chains={a.chains} S={a.S}: best {best:,.0f} mean {a.chains * a.S * a.reps / tot_ms * 1e3:,.0f} steps/s  "
      f"trees/step {c[_abi.RC_CTR_TREES]/steps:.2f} draws/tree {c[_abi.RC_CTR_WALK]/max(c[_abi.RC_CTR_TREES],1):.0f} "
      f"n_sub {c[_abi.RC_CTR_NODES]/steps:.0f} ms {all_ms}", flush=True)
