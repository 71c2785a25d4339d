#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) next to the live CUDA-event breakdown of a
bench.py line:  python scripts/launch_summary.py launches.csv bench.json out.md "title" """
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    launches, bench, out, title = sys.argv[1:5]
    rows = [r for r in csv.reader(open(launches, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
        name = re.sub(r"\(.*", "", r[ki]).strip()
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values()) or 1.0
    line = json.load(open(bench))
    live = line.get("kernel_breakdown_ms_per_step", {})
    live_tot = sum(v[0] for v in live.values()) or 1.0
    live_by = defaultdict(float)
    for k, v in live.items():
        live_by[k.split(":")[0]] += v[0]
    md = [f"# {title}", "", "Per-launch times under ncu are cold-cache and serialised (the 32-NN side stream is serialised too): "
          "compare SHARES with the live CUDA-event breakdown of bench.py, not absolutes.", "",
          "| kernel | launches | total ms | share (ncu) | share (live events, all instantiations) |", "|---|---|---|---|---|"]
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        base = re.sub(r"<.*", "", name).replace("void ", "").replace("poem::", "").strip()
        lv = live_by.get(base)
        md.append(f"| `{name}` | {cnt[name]} | {ms:.3f} | {ms / total:.3f} | {'' if lv is None else f'{lv / live_tot:.3f}'} |")
    md += ["", "Live bench line of the same build (`python bench.py`, not under a profiler):", "", "```json",
           json.dumps({k: line[k] for k in ("value", "ms_per_step", "gpu_launches", "e2e", "roofline", "clocks") if k in line}, indent=1),
           "```", ""]
    open(out, "w").write("\n".join(md))
    print("\n".join(md[:14]))


if __name__ == "__main__":
    main()
