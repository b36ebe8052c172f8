#!/usr/bin/env python
"""bench.py - ReCom steps/s summed over chains (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port)

One bench *step* = one pass of the hot path over the batch: every resident chain advances
by ``--chain-steps`` (default 512) accepted ReCom steps (one launch of the chain-step kernel).  Workload:
BASELINE.json configs[1] (Grid 100x100, 10 districts, eps 0.05, random_spanning_tree ReCom,
1024 chains per GPU); ``--workload config3|config5|config1`` select the others.  Weak
scaling: every rank owns ``--chains`` chains with distinct Philox subsequences; there is
no data-path collective, only the ensemble-histogram all-reduce at
the end of the timed region.

Prints ONE JSON line on rank 0.  `value` = steps/s with chain state resident in HBM;
`e2e` = the same through rc_engine_step_host (pinned host plans in, plans/Tally/cut counts
out, copies inside the timed region); `roofline` = algorithmic HBM bytes of the timed
launches (SURVEY.md section 8d formula, from the kernel's own n_sub/d_sub counters) divided by
the chain-step kernel's CUDA-event time, against MEASURED_PEAKS.json; `cpu_baseline` = the
C oracle port timed on this box's host cores on a bounded sample.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4", "config5"])
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--chain-steps", type=int, default=512, help="ReCom steps per chain per bench step")
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads-per-cta", type=int, default=0)
    ap.add_argument("--wilson", action="store_true", help="uniform_spanning_tree instead of random_spanning_tree")
    ap.add_argument("--cap-nodes", type=int, default=0)
    ap.add_argument("--cap-edges", type=int, default=0)
    return ap.parse_args()


WORKLOAD_NAMES = {
    "config1": "Grid 20x20 (400 nodes), 4 districts, eps=0.05, random_spanning_tree ReCom",
    "config2": "Grid 100x100 (10k nodes), 10 districts, eps=0.05, random_spanning_tree ReCom",
    "config3": "Delaunay 9k nodes, 18 districts, eps=0.02, random_spanning_tree ReCom",
    "config4": "Delaunay 200k nodes, 50 districts, eps=0.05, uniform_spanning_tree (Wilson) ReCom, spill mode",
    "config5": "Delaunay 9k nodes, 18 districts, eps=0.02, region-aware ReCom (67 counties, surcharge 0.5)",
}


def make_workload(args, n_chains, chain_offset=0):
    from gerrychain_b200 import workloads

    fn = getattr(workloads, args.workload)
    kw = dict(n_chains=n_chains, seed=args.seed, chain_offset=chain_offset)
    if args.threads_per_cta:
        kw["threads_per_cta"] = args.threads_per_cta
    if args.wilson:
        from gerrychain_b200 import _abi

        kw["tree_kind"] = _abi.RC_TREE_WILSON
    if args.cap_nodes:
        kw["cap_nodes"] = args.cap_nodes
    if args.cap_edges:
        kw["cap_edges"] = args.cap_edges
    return fn(**kw)


def config_dict(args, world, extra=None):
    d = {
        "workload": WORKLOAD_NAMES[args.workload] + (" [uniform_spanning_tree]" if args.wilson and args.workload != "config4" else ""),
        "chains_per_gpu": args.chains,
        "chains_total": args.chains * world,
        "chain_steps_per_bench_step": args.chain_steps,
        "node_repeats": 1,
        "seed_plan": "10 row stripes" if args.workload == "config2" else
                     ("quadrants" if args.workload == "config1" else "reference recursive_tree_part (committed fixture)"),
        "parallelism": f"chains sharded over {world} GPU(s), no data-path collective",
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


# ---------------------------------------------------------------------------- CPU arm
def cpu_sample(args, seconds, threads=None):
    """Times the C oracle port (oracle/recom_oracle.c) on host cores: `threads` chains at a
    time, enough steps to fill about `seconds`.  Returns (steps/s, cores, description)."""
    from oracle import oracle

    oracle.build()
    cores = threads or os.cpu_count() or 1
    n_chains = cores * 2
    csr, plan, cfg = make_workload(args, n_chains)
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
    """--impl reference: the CPU implementation of the path on this box's host cores.  The
    reference is pure Python and does not travel to the GPU box; the arm therefore times the
    oracle port (C restatement of the reference's algorithm, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle

    oracle.build()
    cores = os.cpu_count() or 1
    n_chains = cores * 2
    csr, plan, cfg = make_workload(args, n_chains)
    oc = oracle.OracleChains(csr, cfg, plan)
    oc.run(2, n_threads=cores)
    t0 = time.perf_counter()
    oc.run(2, n_threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    # each bench step: a bounded sample, about 2 s of CPU work
    per = max(1, int(2.0 / (dt / 2)))
    for _ in range(args.warmup):
        oc.run(per, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oc.run(per, n_threads=cores)
    dt = time.perf_counter() - t0
    assert (oc.status == 0).all()
    value = n_chains * per * args.steps / dt
    sample = f"{n_chains} chains x {per} ReCom steps per bench step on {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
        "data": "synthetic",
        "config": config_dict(args, args.gpus, {"chains_per_gpu": n_chains, "chains_total": n_chains,
                                                "chain_steps_per_bench_step": per, "l2": "n/a (CPU)",
                                                "parallelism": f"{cores} host threads, one chain per thread"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------- GPU arm
def algorithmic_bytes(d_nodes, d_slots, d_steps, a=1):
    """B_step = n_sub*(2a+12) + d_sub*(4+a) + 32, counted once per
    accepted step (the counters count once per proposal; invalid proposals are included,
    retries of the tree are not)."""
    return d_nodes * (2 * a + 12) + d_slots * (4 + a) + 32 * d_steps


def run_b200(args):
    import torch
    import torch.distributed as dist

    from gerrychain_b200 import _abi
    from gerrychain_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])

    csr, plan, cfg = make_workload(args, args.chains, chain_offset=rank * args.chains)
    eng = Engine(csr, cfg, plan, device=local)
    info = eng.info()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    S = args.chain_steps

    eng.run(10).synchronize()  # decorrelate from the seed plan
    eng.check_status()
    for _ in range(args.warmup):
        flush.fill_(1)
        eng.run(S)
    eng.synchronize()

    ctr0 = eng.counters.sum(0).cpu().numpy()
    steps0 = int(eng.step_no.sum().item())
    launches0 = eng.info().kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0.record()
    kernel_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        eng.run(S)
        kernel_ms += eng.last_run_ms()  # CUDA events around the chain-step kernel on its stream
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    eng.check_status()
    launches = eng.info().kernel_launches - launches0
    ctr = eng.counters.sum(0).cpu().numpy() - ctr0
    steps_done = int(eng.step_no.sum().item()) - steps0
    assert steps_done == args.chains * S * args.steps, (steps_done, args.chains, S, args.steps)

    # ensemble histograms: all-reduced over ranks (the only collective of the path)
    cut_hist, dev_hist = eng.histograms(4096, 64, 0.001)
    if world > 1:
        dist.all_reduce(cut_hist)
        dist.all_reduce(dev_hist)
    t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(x) for x in ctr] + [float(steps_done)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms, kernel_ms = float(t[0]), float(t[1])
    tot = tot.cpu().numpy()
    total_steps = tot[-1]
    value = total_steps / (ms * 1e-3)

    # ---- end to end through the host-buffer call ------------------------------------
    h_assign, h_tally, h_cut, h_status = eng.host_buffers()
    h_assign.copy_(eng.assignment[:, : csr.n_nodes].cpu())
    for _ in range(2):
        eng.step_host(h_assign, S, h_assign, h_tally, h_cut, h_status)
    barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(args.steps):
        flush.fill_(1)
        eng.step_host(h_assign, S, h_assign, h_tally, h_cut, h_status)
    ee1.record()
    torch.cuda.synchronize()
    e_wall = time.perf_counter() - e0
    e_ms = max(ee0.elapsed_time(ee1), e_wall * 1e3)
    assert int(h_status.max()) == 0
    te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_ms = float(te[0])
    e2e_value = world * args.chains * S * args.steps / (e_ms * 1e-3)
    h2d = args.chains * csr.n_nodes
    d2h = args.chains * (csr.n_nodes + 8 * cfg.n_districts + 4 + 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    bytes_total = algorithmic_bytes(tot[_abi.RC_CTR_NODES], tot[_abi.RC_CTR_SLOTS], total_steps)
    # per GPU: bytes of this rank's launches / this rank's kernel time == total/world / kernel_ms
    achieved = bytes_total / world / (kernel_ms * 1e-3) / 1e9
    traffic = on_chip = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and not args.wilson:
        try:
            prof = json.load(open(tp)).get(args.workload, {})
            per_step = prof.get("dram_bytes_per_recom_step")
            if per_step is not None:  # per launch, like `achieved`
                traffic = per_step * total_steps / world / args.steps
            on_chip = prof.get("on_chip")  # ncu counters of the same capture: what bounds the kernel
        except Exception:
            traffic = on_chip = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "kernel": "rc::recom_step_kernel",
        "kernel_ms_per_launch": kernel_ms / args.steps,
        "algorithmic_bytes_per_launch": bytes_total / world / args.steps,
        "bytes_per_recom_step": bytes_total / total_steps,
        "trees_per_step": tot[_abi.RC_CTR_TREES] / total_steps,
        "trees_per_sec": tot[_abi.RC_CTR_TREES] / (ms * 1e-3),
        "boruvka_rounds_per_tree": tot[_abi.RC_CTR_ROUNDS] / max(tot[_abi.RC_CTR_TREES], 1),
        "mean_n_sub": tot[_abi.RC_CTR_NODES] / total_steps,
        "mean_d_sub": tot[_abi.RC_CTR_SLOTS] / total_steps,
        "on_chip": on_chip,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": config_dict(args, world, {
            "threads_per_cta": info.threads_per_cta, "ctas_per_sm": info.ctas_per_sm,
            "smem_bytes_per_cta": info.smem_bytes, "cap_nodes": info.cap_nodes, "cap_edges": info.cap_edges}),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e_ms / args.steps, "call": "rc_engine_step_host"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "ensemble": {"mean_cut_edges": float((cut_hist.cpu().numpy() * np.arange(4096)).sum() / max(int(cut_hist.sum()), 1)),
                     "chains_in_histogram": int(cut_hist.sum())},
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_sample(args, args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
