#!/usr/bin/env python
"""Dev tool: local-memory loads/stores (register spills) of the plain and the shared-retry build of the
step kernel by innermost source line, main region only (out-of-line functions excluded).
   python scripts/dev/spill_lines.py [gerrychain_b200/librecom_b200.so]"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "gerrychain_b200/librecom_b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
lines = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout.split("\n")
def analyze(tag, want="main"):
    start = next(i for i, l in enumerate(lines) if l.startswith("//---------------------") and tag in l)
    end = next(i for i in range(start + 10, len(lines)) if lines[i].startswith("//---------------------"))
    region, cnt, cur, ingroup = "main", collections.Counter(), None, False
    for l in lines[start:end]:
        if l.startswith("\t//## File"):
            if not ingroup:
                cur = int(re.search(r"line (\d+)", l).group(1)); ingroup = True
            continue
        ingroup = False
        m = re.match(r"\$_ZN2rc17recom_step_kernel\w+\$_ZN2rc(\d+)(\w+?)IL", l)
        if m: region = m.group(2); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)[\. ]", l)
        if m and m.group(1) in ("LDL", "STL") and region == want: cnt[(m.group(1), cur)] += 1
    return cnt
w = "1" if "--wilson" in sys.argv else "0"
reg = next((x.split("=")[1] for x in sys.argv if x.startswith("--region=")), "main")
a, b = analyze(f"recom_step_kernelILb{w}ELb0ELb0"), analyze(f"recom_step_kernelILb{w}ELb0ELb1", reg)
for k in sorted(set(a) | set(b), key=lambda k: (k[1], k[0])):
    if abs(a[k] - b[k]) >= (1 if "--all" in sys.argv else 3): print(k, "plain", a[k], "share", b[k])
print("total", sum(a.values()), sum(b.values()))
