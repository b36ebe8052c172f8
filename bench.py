#!/usr/bin/env python
"""bench.py - ReCom steps/s summed over chains (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU ReCom

One bench *step* = one pass of the hot path over the batch: every resident chain advances
by ``--chain-steps`` (default 512) accepted ReCom steps (one launch of the chain-step
kernel), followed by the ensemble-histogram kernel and - at N > 1 - its NCCL all-reduce
(the one collective of the path).  Headline workload: BASELINE.json
configs[1] (Grid 100x100, 10 districts, eps 0.05, random_spanning_tree ReCom, 1024 chains
per GPU; weak scaling).  The same process then times the other BASELINE configurations
and reports them under ``extra`` (config 3 = 1024 chains per GPU, i.e. 8192 on 8 GPUs;
config 5; config 3 at a fixed total of 8192 chains = the strong-scaling figure; config 3
with uniform_spanning_tree; config 4).

Prints ONE JSON line on rank 0.  `value` = steps/s with chain state resident in HBM;
`e2e` = the same through rc_engine_step_host (pinned host plans in, plans/Tally/cut counts
out, copies inside the timed region; the copies are amortised over ``--chain-steps`` steps
per chain, which is how a batched chain runner is used - one plan per chain per
checkpoint); `roofline` = algorithmic HBM bytes of the timed launches (SURVEY.md
section 8d formula, from the kernel's own n_sub/d_sub counters) divided by the chain-step
kernel's CUDA-event time, against MEASURED_PEAKS.json; `cpu_baseline` = the UNMODIFIED
reference (mggg/GerryChain, oracle/_ref, one chain process per host core) on a bounded
sample, `cpu_baseline_port` = the C oracle port on all host threads.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "recom_steps_per_sec"
UNIT = "steps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_SHIM = os.path.join(ROOT, "oracle", "refshim")
WORKLOADS = ["config1", "config2", "config3", "config4", "config5"]
N_CUT_BINS = 1 << 15  # cut-edge histogram bins (config 4 plans have ~12k cut edges)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=WORKLOADS)
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--chain-steps", type=int, default=512, help="ReCom steps per chain per bench step")
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of each cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"],
                    help="--impl reference: the Python reference (oracle/_ref) or the C oracle port")
    ap.add_argument("--extras", default="default",
                    help="comma list of extra workloads timed in the same process "
                         "(config3,config5,config3_strong,config3_short,config3_wilson,config4), 'default' or 'none'")
    ap.add_argument("--threads-per-cta", type=int, default=0)
    ap.add_argument("--wilson", action="store_true", help="uniform_spanning_tree instead of random_spanning_tree")
    ap.add_argument("--cap-nodes", type=int, default=0)
    ap.add_argument("--cap-edges", type=int, default=0)
    return ap.parse_args()


WORKLOAD_NAMES = {
    "config1": "Grid 20x20 (400 nodes), 4 districts, eps=0.05, random_spanning_tree ReCom",
    "config2": "Grid 100x100 (10k nodes), 10 districts, eps=0.05, random_spanning_tree ReCom",
    "config3": "Delaunay 9k nodes, 18 districts, eps=0.02, random_spanning_tree ReCom",
    "config4": "Delaunay 200k nodes, 50 districts, eps=0.05, uniform_spanning_tree (Wilson) ReCom",
    "config5": "Delaunay 9k nodes, 18 districts, eps=0.02, region-aware ReCom (67 counties, surcharge 0.5)",
}


def workload_label(name, wilson):
    return WORKLOAD_NAMES[name] + (" [uniform_spanning_tree]" if wilson and name != "config4" else "")


def make_workload(name, n_chains, seed, chain_offset=0, wilson=False, threads_per_cta=0, cap_nodes=0, cap_edges=0):
    from gerrychain_b200 import workloads

    kw = dict(n_chains=n_chains, seed=seed, chain_offset=chain_offset)
    if threads_per_cta:
        kw["threads_per_cta"] = threads_per_cta
    if wilson:
        from gerrychain_b200 import _abi

        kw["tree_kind"] = _abi.RC_TREE_WILSON
    if cap_nodes:
        kw["cap_nodes"] = cap_nodes
    if cap_edges:
        kw["cap_edges"] = cap_edges
    return getattr(workloads, name)(**kw)


def seed_plan_label(name):
    return {"config2": "10 row stripes", "config1": "quadrants"}.get(
        name, "reference recursive_tree_part (committed fixture)")


def config_dict(workload, wilson, chains, world, chain_steps, extra=None):
    d = {
        "workload": workload_label(workload, wilson),
        "chains_per_gpu": chains,
        "chains_total": chains * world,
        "chain_steps_per_bench_step": chain_steps,
        "node_repeats": 1,
        "seed_plan": seed_plan_label(workload),
        "parallelism": f"chains sharded over {world} GPU(s), no data-path collective; "
                       "ensemble histograms all-reduced (NCCL) every bench step",
        "l2": "256 MiB flush buffer written between timed steps (inside the timed region)",
    }
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------- CPU arms
def have_reference():
    return os.path.isdir(os.path.join(REF_DIR, "gerrychain"))


def _ref_worker(conn, workload, wilson, worker_id):
    """One chain of the UNMODIFIED reference in its own process (the reference's own way of
    running ensembles, /root/reference/tests/test_region_aware.py:82-99): the loop timed is
    /root/reference/gerrychain/chain.py:166-196 with proposals/tree_proposals.py:36-146."""
    import random
    import warnings
    from functools import partial

    warnings.filterwarnings("ignore")
    sys.path[:0] = [REF_DIR, REF_SHIM]
    from gerrychain import Graph, MarkovChain, Partition, accept, constraints
    from gerrychain.proposals import recom
    from gerrychain.tree import bipartition_tree, uniform_spanning_tree
    from gerrychain.updaters import Tally, cut_edges

    from gerrychain_b200 import synthetic as syn

    golden = os.path.join(ROOT, "tests", "golden")
    regions = None
    if workload in ("config1", "config2"):
        m, k = (20, 4) if workload == "config1" else (100, 10)
        g = syn.grid_graph(m, m)
        plan = syn.grid_quadrants(m, m) if k == 4 else syn.grid_stripes(m, m, k)
        eps = 0.05
    else:
        n, k, eps, pr = (200000, 50, 0.05, (20, 200)) if workload == "config4" else (9000, 18, 0.02, (500, 2500))
        g = syn.delaunay_graph(n, seed=0, pop_range=pr)
        if workload == "config5":
            syn.voronoi_counties(g, 67, seed=1, key="county")
            regions = {"county": 0.5}
        vec = np.load(os.path.join(golden, f"plan_delaunay{n}_k{k}.npy"))
        plan = {v: int(d) for v, d in zip(g.nodes, vec)}
        wilson = wilson or workload == "config4"
    graph = Graph.from_networkx(g)
    part = Partition(graph, plan, {"population": Tally("population"), "cut_edges": cut_edges})
    ideal = sum(part["population"].values()) / len(part)
    method = partial(bipartition_tree, spanning_tree_fn=uniform_spanning_tree) if wilson else bipartition_tree
    kw = dict(pop_col="population", pop_target=ideal, epsilon=eps, node_repeats=1, method=method)
    if regions:
        kw["region_surcharge"] = regions
    random.seed(worker_id)
    chain = iter(MarkovChain(partial(recom, **kw), [constraints.within_percent_of_ideal_population(part, eps)],
                             accept.always_accept, part, total_steps=1 << 60))
    next(chain)  # the initial state
    conn.send(("ready", 0.0))
    while True:
        n_steps = conn.recv()
        if n_steps <= 0:
            break
        t0 = time.perf_counter()
        touched = 0
        for _ in range(n_steps):
            state = next(chain)
            touched += len(state["cut_edges"])
        conn.send(("done", time.perf_counter() - t0))


class ReferencePool:
    """`cores` reference chains, one process each, advanced in lock step."""

    def __init__(self, workload, wilson, cores):
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        self.procs, self.conns = [], []
        for w in range(cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_worker, args=(b, workload, wilson, w), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        for c in self.conns:
            c.recv()

    def run(self, n_steps):
        """every chain advances n_steps; returns max worker wall time"""
        for c in self.conns:
            c.send(n_steps)
        return max(c.recv()[1] for c in self.conns)

    def close(self):
        for c in self.conns:
            try:
                c.send(0)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def port_sample(workload, wilson, seed, seconds, threads=None):
    """Times the C oracle port (oracle/recom_oracle.c) on host cores: `threads` chains at a
    time, enough steps to fill about `seconds`.  Returns (steps/s, cores, description)."""
    from oracle import oracle

    oracle.build()
    cores = threads or os.cpu_count() or 1
    n_chains = cores * 2
    csr, plan, cfg = make_workload(workload, n_chains, seed, wilson=wilson)
    oc = oracle.OracleChains(csr, cfg, plan)
    oc.run(2, n_threads=cores)  # warm-up + calibration
    t0 = time.perf_counter()
    oc.run(2, n_threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    per = max(2, int(seconds / (dt / 2)))
    per = min(per, 100000)
    t0 = time.perf_counter()
    oc.run(per, n_threads=cores)
    dt = time.perf_counter() - t0
    assert (oc.status == 0).all(), "oracle chain stopped"
    return n_chains * per / dt, cores, f"{n_chains} chains x {per} steps on {cores} host threads ({dt:.1f} s)"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's
    host cores - the unmodified Python package from oracle/_ref (one chain process per core);
    the C oracle port only when oracle/_ref is absent or --ref-kind port is given."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = args.ref_kind
    if kind == "auto":
        kind = "reference" if have_reference() else "port"
    if kind == "reference" and not have_reference():
        raise SystemExit("bench.py: oracle/_ref is missing (sh oracle/make_ref.sh in the build container)")
    wilson = args.wilson or args.workload == "config4"
    if kind == "reference":
        pool = ReferencePool(args.workload, wilson, cores)
        dt = pool.run(1)
        dt = max(pool.run(2) / 2, 1e-3, dt / 4)  # calibration: seconds per ReCom step per chain
        per = max(1, int(2.0 / dt))  # each bench step: a bounded sample, about 2 s of CPU work
        for _ in range(args.warmup):
            pool.run(per)
        wall = 0.0
        for _ in range(args.steps):
            wall += pool.run(per)  # sum over bench steps of the slowest worker's time
        pool.close()
        n_chains = cores
        sample = (f"unmodified mggg/GerryChain (oracle/_ref): {cores} chain processes x {per} ReCom steps per bench "
                  f"step, one per host core, random.seed(worker)")
        par = f"{cores} host processes, one reference chain per core"
    else:
        from oracle import oracle

        oracle.build()
        n_chains = cores * 2
        csr, plan, cfg = make_workload(args.workload, n_chains, args.seed, wilson=args.wilson)
        oc = oracle.OracleChains(csr, cfg, plan)
        oc.run(2, n_threads=cores)
        t0 = time.perf_counter()
        oc.run(2, n_threads=cores)
        dt = max(time.perf_counter() - t0, 1e-4)
        per = max(1, int(2.0 / (dt / 2)))
        for _ in range(args.warmup):
            oc.run(per, n_threads=cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oc.run(per, n_threads=cores)
        wall = time.perf_counter() - t0
        assert (oc.status == 0).all()
        sample = f"C oracle port: {n_chains} chains x {per} ReCom steps per bench step on {cores} host threads"
        par = f"{cores} host threads, one chain per thread"
    value = n_chains * per * args.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
        "data": "synthetic",
        "config": config_dict(args.workload, args.wilson, n_chains, 1, per,
                              {"chains_total": n_chains, "l2": "n/a (CPU)", "parallelism": par}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def reference_sample(workload, wilson, seconds):
    """cpu_baseline of the GPU arm: the reference arm as a child process (no fork after CUDA
    initialisation), sized to about `seconds` of timed CPU work."""
    steps = max(2, int(round(seconds / 2.0)))
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-kind", "reference",
           "--workload", workload, "--steps", str(steps), "--warmup", "1"] + (["--wilson"] if wilson else [])
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600, env=env)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return None


# ---------------------------------------------------------------------------- GPU arm
def algorithmic_bytes(d_nodes, d_slots, d_steps, a=1):
    """B_step = n_sub*(2a+12) + d_sub*(4+a) + 32, counted once per
    accepted step (the counters count once per proposal; invalid proposals are included,
    retries of the tree are not)."""
    return d_nodes * (2 * a + 12) + d_slots * (4 + a) + 32 * d_steps


class Dist:
    def __init__(self):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])

    def all_reduce(self, t, op=None):
        if self.world > 1:
            self.dist.all_reduce(t, op=op if op is not None else self.dist.ReduceOp.SUM)
        return t

    def max_(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(D, flush, workload, chains, S, steps, warmup, seed, wilson=False, e2e=False, sampler=None,
            threads_per_cta=0, cap_nodes=0, cap_edges=0):
    """Times `steps` bench steps of `workload` with `chains` chains on this rank; returns the
    whole-job numbers (all ranks) as a dict.  Every rank calls it with the same arguments."""
    import torch

    from gerrychain_b200 import _abi
    from gerrychain_b200.engine import Engine

    csr, plan, cfg = make_workload(workload, chains, seed, chain_offset=D.rank * chains, wilson=wilson,
                                   threads_per_cta=threads_per_cta, cap_nodes=cap_nodes, cap_edges=cap_edges)
    eng = Engine(csr, cfg, plan, device=D.local)
    info = eng.info()
    cut_hist = dev_hist = None

    def bench_step():
        nonlocal cut_hist, dev_hist
        flush.fill_(1)
        eng.run(S)
        ms_k = eng.last_run_ms()  # CUDA events around the chain-step kernel on its stream
        cut_hist, dev_hist = eng.histograms(N_CUT_BINS, 64, 0.001)  # the checkpoint: ensemble histograms ...
        D.all_reduce(cut_hist)                                 # ... all-reduced over the ranks (NCCL)
        D.all_reduce(dev_hist)
        return ms_k

    eng.run(10).synchronize()  # decorrelate from the seed plan
    eng.check_status()
    for _ in range(warmup):
        bench_step()
    eng.synchronize()
    ctr0 = eng.counters.sum(0).cpu().numpy()
    steps0 = int(eng.step_no.sum().item())
    launches0 = eng.info().kernel_launches
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0.record()
    kernel_ms = 0.0
    for _ in range(steps):
        kernel_ms += bench_step()
    ev1.record()
    torch.cuda.synchronize()
    D.barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None
    eng.check_status()
    launches = eng.info().kernel_launches - launches0
    ctr = eng.counters.sum(0).cpu().numpy() - ctr0
    steps_done = int(eng.step_no.sum().item()) - steps0
    assert steps_done == chains * S * steps, (steps_done, chains, S, steps)
    t = D.max_(torch.tensor([ms, kernel_ms], dtype=torch.float64, device=D.dev))
    tot = D.all_reduce(torch.tensor([float(x) for x in ctr] + [float(steps_done)], dtype=torch.float64, device=D.dev))
    ms, kernel_ms = float(t[0]), float(t[1])
    tot = tot.cpu().numpy()
    total_steps = tot[-1]
    bytes_total = algorithmic_bytes(tot[_abi.RC_CTR_NODES], tot[_abi.RC_CTR_SLOTS], total_steps)
    out = {
        "value": total_steps / (ms * 1e-3), "ms": ms, "kernel_ms": kernel_ms, "total_steps": total_steps,
        "launches": int(launches), "clocks": clocks, "info": info, "csr": csr, "cfg": cfg,
        "bytes_total": bytes_total, "tot": tot,
        # per GPU: bytes of this rank's launches / this rank's kernel time == total/world / kernel_ms
        "achieved_gbs": bytes_total / D.world / (kernel_ms * 1e-3) / 1e9,
        "mean_cut_edges": float((cut_hist.cpu().numpy() * np.arange(N_CUT_BINS)).sum() / max(int(cut_hist.sum()), 1)),
        "chains_in_histogram": int(cut_hist.sum()),
    }
    if e2e:
        # ---- end to end through the host-buffer call ------------------------------------
        h_assign, h_tally, h_cut, h_status = eng.host_buffers()
        h_assign.copy_(eng.assignment[:, : csr.n_nodes].cpu())
        for _ in range(2):
            eng.step_host(h_assign, S, h_assign, h_tally, h_cut, h_status)
        D.barrier()
        torch.cuda.synchronize()
        e0 = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(steps):
            flush.fill_(1)
            eng.step_host(h_assign, S, h_assign, h_tally, h_cut, h_status)
        ee1.record()
        torch.cuda.synchronize()
        e_wall = time.perf_counter() - e0
        e_ms = max(ee0.elapsed_time(ee1), e_wall * 1e3)
        assert int(h_status.max()) == 0
        e_ms = float(D.max_(torch.tensor([e_ms], dtype=torch.float64, device=D.dev))[0])
        out["e2e"] = {"value": D.world * chains * S * steps / (e_ms * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": chains * csr.n_nodes,
                      "d2h_bytes_per_step": chains * (csr.n_nodes + 8 * cfg.n_districts + 4 + 4),
                      "ms_per_step": e_ms / steps, "call": "rc_engine_step_host",
                      "note": f"host plans in/out once per {S} ReCom steps per chain (one checkpoint)"}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return out


def roofline_dict(m, steps, world, peak, peak_src, workload, wilson):
    from gerrychain_b200 import _abi

    tot, total_steps = m["tot"], m["total_steps"]
    traffic = on_chip = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    key = workload + ("_wilson" if wilson and workload != "config4" else "")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp)).get(key, {})
            per_step = prof.get("dram_bytes_per_recom_step")
            if per_step is not None:  # per launch, like `achieved`
                traffic = per_step * total_steps / world / steps
            on_chip = prof.get("on_chip")  # ncu counters of the same capture: what bounds the kernel
        except Exception:
            traffic = on_chip = None
    return {
        "bound": "hbm", "achieved": m["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": m["achieved_gbs"] / peak,
        "traffic": traffic, "peak_source": peak_src, "kernel": "rc::recom_step_kernel",
        "kernel_ms_per_launch": m["kernel_ms"] / steps,
        "algorithmic_bytes_per_launch": m["bytes_total"] / world / steps,
        "bytes_per_recom_step": m["bytes_total"] / total_steps,
        "trees_per_step": tot[_abi.RC_CTR_TREES] / total_steps,
        "trees_per_sec": tot[_abi.RC_CTR_TREES] / (m["ms"] * 1e-3),
        "boruvka_rounds_per_tree": tot[_abi.RC_CTR_ROUNDS] / max(tot[_abi.RC_CTR_TREES], 1),
        "mean_n_sub": tot[_abi.RC_CTR_NODES] / total_steps,
        "mean_d_sub": tot[_abi.RC_CTR_SLOTS] / total_steps,
        "on_chip": on_chip,
    }


# name -> (workload, wilson, chains per GPU as f(world), chain steps per bench step, bench steps, scaling)
EXTRAS = {
    "config3": ("config3", False, lambda w: 1024, 512, 5, "weak"),
    "config5": ("config5", False, lambda w: 1024, 512, 5, "weak"),
    "config3_strong": ("config3", False, lambda w: 8192 // w, 64, 5, "strong"),
    # short launches (a checkpoint every 32 steps): the heavy-tailed retries of a few chains set
    # the pace unless the CTAs that run out of chains help (shared retries)
    "config3_short": ("config3", False, lambda w: 1024, 32, 8, "weak"),
    "config3_wilson": ("config3", True, lambda w: 1024, 256, 4, "weak"),
    "config4": ("config4", True, lambda w: 1024, 8, 3, "weak"),
}
DEFAULT_EXTRAS = ["config3", "config5", "config3_strong", "config3_short", "config3_wilson", "config4"]


def run_b200(args):
    import torch

    D = Dist()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=D.dev)
    S = args.chain_steps
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    m = measure(D, flush, args.workload, args.chains, S, args.steps, args.warmup, args.seed, wilson=args.wilson,
                e2e=True, sampler=sampler, threads_per_cta=args.threads_per_cta, cap_nodes=args.cap_nodes,
                cap_edges=args.cap_edges)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"

    names = [] if args.extras == "none" else (DEFAULT_EXTRAS if args.extras == "default" else args.extras.split(","))
    extra = {}
    for name in names:
        wl, wilson, chains_of, cs, bsteps, scaling = EXTRAS[name]
        chains = chains_of(D.world)
        try:
            x = measure(D, flush, wl, chains, cs, bsteps, 3, args.seed, wilson=wilson)
        except Exception as ex:  # an extra must never take the headline line down with it
            extra[name] = {"error": f"{type(ex).__name__}: {ex}"}
            continue
        r = roofline_dict(x, bsteps, D.world, peak, peak_src, wl, wilson)
        extra[name] = {
            "workload": workload_label(wl, wilson), "value": x["value"], "unit": UNIT, "scaling": scaling,
            "chains_per_gpu": chains, "chains_total": chains * D.world, "chain_steps_per_bench_step": cs,
            "steps": bsteps, "warmup": 3, "ms_per_step": x["ms"] / bsteps, "gpu_launches": x["launches"],
            "trees_per_step": r["trees_per_step"], "roofline_frac": r["frac"], "achieved_gbs": r["achieved"],
            "bytes_per_recom_step": r["bytes_per_recom_step"], "mean_cut_edges": x["mean_cut_edges"],
            "threads_per_cta": x["info"].threads_per_cta, "ctas_per_sm": x["info"].ctas_per_sm,
        }
    if D.rank != 0:
        D.close()
        return

    info, csr, cfg = m["info"], m["csr"], m["cfg"]
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": config_dict(args.workload, args.wilson, args.chains, D.world, S, {
            "threads_per_cta": info.threads_per_cta, "ctas_per_sm": info.ctas_per_sm,
            "smem_bytes_per_cta": info.smem_bytes, "cap_nodes": info.cap_nodes, "cap_edges": info.cap_edges}),
        "clocks": m["clocks"],
        "e2e": m["e2e"],
        "gpu_launches": m["launches"],
        "roofline": roofline_dict(m, args.steps, D.world, peak, peak_src, args.workload, args.wilson),
        "ensemble": {"mean_cut_edges": m["mean_cut_edges"], "chains_in_histogram": m["chains_in_histogram"]},
        "extra": extra,
    }
    line["cpu_baseline"] = line["cpu_baseline_port"] = None
    if not args.no_cpu_baseline and D.world == 1:
        wilson = args.wilson or args.workload == "config4"
        v, cores, sample = port_sample(args.workload, args.wilson, args.seed, args.cpu_seconds)
        line["cpu_baseline_port"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        ref = None
        if have_reference():
            try:
                ref = reference_sample(args.workload, wilson, args.cpu_seconds)
            except Exception:
                ref = None
        line["cpu_baseline"] = ref if ref else line["cpu_baseline_port"]
        # the north star quotes its target ratio on the 9k-node config: the unmodified reference
        # timed there too (same box, same run), so that extra.config3.value / this is that ratio
        if have_reference() and "value" in extra.get("config3", {}):
            try:
                extra["config3"]["cpu_baseline"] = reference_sample("config3", False, min(args.cpu_seconds, 10.0))
            except Exception:
                extra["config3"]["cpu_baseline"] = None
    emit(line)
    D.close()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line, on the real stdout (everything else a library prints - e.g. NCCL's
    version banner - has been diverted to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stdout of this process and of the libraries it loads -> stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
