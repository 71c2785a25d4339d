#!/usr/bin/env python
"""Summarise an .ncu-rep (every kernel launch in it) into a short markdown table: python scripts/ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
]


def raw_all(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, vals)} for vals in rows[2:]]


def main():
    rep = sys.argv[1]
    text = "\n".join(one(rep, m) for m in raw_all(rep))
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


def one(rep, m):
    lines = [f"# ncu --set full: `{m.get('Kernel Name', ('?', ''))[0][:90]}`", "", f"report: `{rep}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    stalls = sorted(((float(v[0].replace(',', '')), k) for k, v in m.items()
                     if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                     and v[0] not in ("", "n/a")), reverse=True)[:8]
    lines += ["", "Top warp stall reasons (avg warps stalled per issue-active cycle):", ""]
    for val, k in stalls:
        lines.append(f"- {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {val:.2f}")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    main()
