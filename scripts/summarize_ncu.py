#!/usr/bin/env python
"""Turn an ncu report (gpurun_out/*.ncu-rep) into the text summary committed under profiles/:
launch configuration, the counters the north star names, stall reasons, and the per-function /
per-source-line distribution of samples and instructions (scripts/ncu_lines.py).
usage: summarize_ncu.py report.ncu-rep "title" steps_in_launch > profiles/xyz.txt"""
import csv, subprocess, sys, io, os
rep, title, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], zip(rows[2], rows[1])))
def g(k):
    return float(d[k][0].replace(",", "")) if k in d and d[k][0] not in ("", "n/a") else float("nan")
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"{title}\nncu --set full --clock-control none --import-source on, one launch" + (f"; built from commit {head} (+ working tree)" if head else "") + "\n")
inst = g("smsp__inst_executed.sum")
print(f"launch: grid {int(g('launch__grid_size'))} x {int(g('launch__block_size'))} threads, {int(g('launch__registers_per_thread'))} registers/thread, "
      f"{g('launch__shared_mem_per_block_dynamic'):.1f} KB dynamic shared memory per CTA; CTAs/SM limited by registers {int(g('launch__occupancy_limit_registers'))}, "
      f"shared memory {int(g('launch__occupancy_limit_shared_mem'))}, warps {int(g('launch__occupancy_limit_warps'))}")
print(f"duration {g('gpu__time_duration.sum'):.2f} ms for {int(steps)} ReCom steps")
print(f"warp instructions {inst:.4g}  = {inst / steps / 1e3:.0f} k per ReCom step; threads active per warp instruction {g('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} of 32")
for k, name in [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (% of peak)"),
                ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64 per SM)"),
                ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
                ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (% of peak)"),
                ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
                ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)")]:
    if k in d:
        print(f"{name:45s} {g(k):8.2f}")
dr, dw = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
ur, uw = d.get("dram__bytes_read.sum", ("", ""))[1], d.get("dram__bytes_write.sum", ("", ""))[1]
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
tot = dr * mult.get(ur, 1.0) + dw * mult.get(uw, 1.0)
print(f"DRAM traffic {tot / 1e6:.1f} MB per launch = {tot / steps:.0f} B per ReCom step (read {dr:.1f} {ur}, write {dw:.1f} {uw})")
sw, bc = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
print(f"shared-memory wavefronts {sw:.4g}, of them bank conflicts {bc:.4g} ({100 * bc / sw:.0f} %)")
print("\nstall reasons (warps stalled per issue-active cycle):")
st = [(k.split("issue_stalled_")[1].split("_per_issue")[0], g(k)) for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for name, v in sorted(st, key=lambda x: -x[1])[:9]:
    print(f"  {name:28s} {v:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
tmp = "/tmp/_ncu_src.csv"
open(tmp, "w").write(src)
print("\nper source line (samples / warp instructions) and per function:")
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"), tmp, "28"], capture_output=True, text=True).stdout)
