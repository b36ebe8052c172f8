#!/usr/bin/env python
"""Dev tool: per-phase cycle breakdown of the chain-step kernel from an RC_TIMERS build.

    nvcc ... -DRC_TIMERS -o variants/lib_timers.so   (scripts/build_timers.sh)
    RECOM_B200_LIB=variants/lib_timers.so python scripts/phases.py --workload config2
"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gerrychain_b200 import workloads, _abi
from gerrychain_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config2")
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--chain-steps", type=int, default=8)
ap.add_argument("--cap-nodes", type=int, default=0)
ap.add_argument("--cap-edges", type=int, default=0)
ap.add_argument("--wilson", action="store_true")
ap.add_argument("--seed", type=int, default=5)
ap.add_argument("--walkers", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
a = ap.parse_args()
kw = dict(n_chains=a.chains, seed=a.seed, cap_nodes=a.cap_nodes, cap_edges=a.cap_edges)
if a.walkers:
    kw["wilson_walkers"] = a.walkers
if a.threads:
    kw["threads_per_cta"] = a.threads
if a.wilson:
    kw["tree_kind"] = _abi.RC_TREE_WILSON
csr, plan, cfg = getattr(workloads, a.workload)(**kw)
eng = Engine(csr, cfg, plan)
eng.run(4).synchronize()
ctr0 = eng.counters.sum(0).cpu().numpy()
step0 = eng.step_no.sum().item()
eng.run(a.chain_steps).synchronize()
ms = eng.last_run_ms()
out = (C.c_ulonglong * 16)()
eng.lib.rc_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
eng.lib.rc_debug_phase_cycles(eng._handle, out)
names = ["to-pick", "pick", "gather", "tree", "euler", "balance", "choose", "commit", "flush+fence", "ticket+wait", "state loads", "wilson block rounds+rest", "wilson warp-0 detect", "wilson warp-0 pop"]
tot = sum(out[:14]) or 1
info = eng.info()
steps = a.chains * a.chain_steps
ctr = eng.counters.sum(0).cpu().numpy() - ctr0
if a.wilson:
    print("  walk loop: %.1f cycles per trip, %d trips per batch" % (out[11] / max(out[12], 1), out[12] / max(ctr[7], 1)))
    print("  walk batches %d, tree-phase cycles per batch %.0f, per draw of one lane %.1f" % (ctr[7], out[3] / max(ctr[7], 1), out[3] / max(ctr[7], 1) / max(ctr[6] / max(ctr[0], 1), 1)))
print("  trees/step %.2f  walk draws/tree %.0f  mean n_sub %.0f" % (ctr[0] / max(eng.step_no.sum().item() - step0, 1), ctr[6] / max(ctr[0], 1), ctr[2] / max(eng.step_no.sum().item() - step0, 1)))
print(f"{a.workload} wilson={a.wilson} chains={a.chains} steps/chain={a.chain_steps}: {ms:.2f} ms, "
      f"{steps / ms * 1e3:.0f} steps/s, smem {info.smem_bytes} B, ctas/sm {info.ctas_per_sm}, grid {info.grid}")
names.append("wilson global detections" if a.wilson else "slot14")
for n, v in zip(names, out[:15]):
    print(f"  {n:14s} {100 * v / tot:5.1f}%  {v / steps:12.0f} cycles/step (thread 0 of one CTA)")
