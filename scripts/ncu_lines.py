#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":  # a source line summary row
        d = dict(zip(hdr[4:], r[4:]))
        try:
            out.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]),
                        int(d["Thread Instructions Executed"]), d))
        except (ValueError, KeyError):
            pass
tot_s = sum(o[3] for o in out) or 1; tot_i = sum(o[4] for o in out) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}, thread instructions {sum(o[5] for o in out)}")
out.sort(key=lambda o: -o[3])
for f, ln, src, smp, ins, tins, d in out[:top]:
    stalls = sorted(((k, int(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{100*smp/tot_s:5.1f}% smp {100*ins/tot_i:5.1f}% ins  {f}:{ln:<4} {src}   {stalls}")

# ---- phase table for recom_kernels.cu (function line ranges are read from the file) ----
import os, re
src_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gerrychain_b200", "csrc", "recom_kernels.cu")
if os.path.exists(src_path):
    lines = open(src_path).read().split("\n")
    starts = [(i + 1, m.group(1)) for i, l in enumerate(lines)
              for m in [re.match(r"^(?:template.*\n)?__(?:device|global)__ .*?(\w+)\(", l)] if m]
    def phase(f, ln):
        if f != "recom_kernels.cu": return f
        name = "?"
        for s, nm in starts:
            if s <= ln: name = nm
        return name
    agg = {}
    for f, ln, src, smp, ins, tins, d in out:
        k = phase(f, ln)
        a = agg.setdefault(k, [0, 0]); a[0] += smp; a[1] += ins
    print("\nby function:")
    for k, (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {100*smp/tot_s:5.1f}% smp {100*ins/tot_i:5.1f}% ins  {k}")
