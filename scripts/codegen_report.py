#!/usr/bin/env python
"""Static code-generation report of the CUDA library (no GPU needed): ptxas -v (registers, stack,
spills per kernel and out-of-line function) and a SASS opcode histogram per function.
usage: python scripts/codegen_report.py [8|16] > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gerrychain_b200", "csrc")
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 8
so = f"/tmp/_codegen_{bits}.so"
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
       "-shared", f"-DRC_LABEL_BITS={bits}", "-Xptxas", "-v", "-o", so, "recom_abi.cu"]
out = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


print(f"ptxas -v (nvcc 12.9, -O3, sm_100a) of gerrychain_b200/csrc/recom_abi.cu at {head}, -DRC_LABEL_BITS={bits}: "
      "registers, stack, spills per kernel and per out-of-line function")
print("template arguments of recom_step_kernel: <kWilson, kSpill, kShare>\n")
cur = None
for ln in out.stderr.splitlines():
    m = re.search(r"(?:Compiling entry function|Function properties for) '?([_A-Za-z0-9]+)'?", ln)
    if m:
        cur = demangle(m.group(1))
        continue
    if cur and ("bytes stack frame" in ln or "Used " in ln):
        txt = ln.split("ptxas info    :")[-1].strip() if "ptxas info" in ln else ln.strip()
        if "bytes stack frame" in ln and txt.startswith("0 bytes stack frame, 0 bytes spill stores"):
            continue
        print(f"{cur[:110]:110s} | {txt}")
print("\nSASS opcode histogram per function (cuobjdump -sass, static counts), sm_100a\n")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, hist = None, collections.Counter()


def flush():
    if fn and sum(hist.values()) > 300:
        print(f"{demangle(fn)[:120]}: {sum(hist.values())} instructions")
        print("   " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(28)))


for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        flush()
        fn, hist = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        hist[m.group(1).split(".")[0]] += 1
flush()
os.remove(so)
