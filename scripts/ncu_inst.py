#!/usr/bin/env python
"""Executed warp instructions and shared-memory wavefronts per source line of one kernel (same join as ncu_lines.py).
   python scripts/ncu_inst.py gpurun_out/x.ncu-rep va_fused_kernelILi256E [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

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
ie, wf, wfi = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ie] or 0), int(r[wf] or 0), int(r[wfi] or 0)))
    except (ValueError, IndexError):
        continue
assert len(data) == len(ins), (len(data), len(ins))
agg = defaultdict(lambda: [0, 0, 0])
ops = defaultdict(lambda: defaultdict(int))
for (n, w, wi), (loc, t) in zip(data, ins):
    a = agg[loc]
    a[0] += n
    a[1] += w
    a[2] += wi
    ops[loc][t.split()[0].split(".")[0]] += n
tot = sum(a[0] for a in agg.values())
totw = sum(a[1] for a in agg.values())
print(f"{func}: {tot} warp instructions, {totw} shared wavefronts")
srcs = {}
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    line = ""
    if loc and loc[0].endswith((".cuh", ".cu")):
        path = os.path.join(ROOT, "poem-v2_b200", "csrc", loc[0])
        if path not in srcs and os.path.exists(path):
            srcs[path] = open(path).read().split("\n")
        if path in srcs:
            line = srcs[path][loc[1] - 1].strip()[:70]
    o = sorted(ops[loc].items(), key=lambda kv: -kv[1])[:3]
    print(f"{str(loc):30s} {100 * a[0] / tot:5.1f}% inst  wavefronts {a[1]:9d} (ideal {a[2]:9d})  {[k for k, _ in o]}  {line}")
