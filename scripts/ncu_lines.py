#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins the SASS page of an .ncu-rep with `nvdisasm -g` of the built library.
   python scripts/ncu_lines.py gpurun_out/va_fused.ncu-rep va_fused_kernelILi256E [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, func = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lib = os.path.join(ROOT, "poem-v2_b200", "csrc", "libpoem_b200.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(sass) if l.startswith(".text.")]
s0 = [i for i in starts if func in sass[i]][0]
s1 = min([i for i in starts if i > s0] + [len(sass)])
cur, ins = None, []
for l in sass[s0:s1]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
si, ns = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= ns:
        continue
    try:
        n = int(r[ns])
    except ValueError:
        continue
    data.append((n, r[si], {h[6:]: int(r[i] or 0) for i, h in stall_cols}))
assert len(data) == len(ins), (len(data), len(ins), "library and report are different builds")
agg, why = defaultdict(int), defaultdict(lambda: defaultdict(int))
for (n, s, st), (loc, t) in zip(data, ins):
    agg[loc] += n
    for k, v in st.items():
        why[loc][k] += v
tot = sum(agg.values())
srcs = {}
print(f"{func}: {tot} samples, {len(ins)} instructions")
for loc, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    line = ""
    if loc and loc[0].endswith((".cuh", ".cu")):
        path = os.path.join(ROOT, "poem-v2_b200", "csrc", loc[0])
        if path not in srcs and os.path.exists(path):
            srcs[path] = open(path).read().split("\n")
        if path in srcs:
            line = srcs[path][loc[1] - 1].strip()[:80]
    w = sorted(why[loc].items(), key=lambda kv: -kv[1])[:2]
    print(f"{str(loc):30s} {n:6d} {100 * n / tot:5.1f}%  {[(k, v) for k, v in w if v]}  {line}")
