#!/usr/bin/env python
"""Summarise an .ncu-rep for profiles/: headline raw metrics + the hottest source lines.

usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/rNN_<kernel>.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = ncu(["-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name"), "| id", d.get("ID"))
        for k in RAW:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                v = float(d[k] or 0)
                if v >= 0.3:
                    print("  stall %-64s %.2f" % (k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
    # per source line: executed warp instructions and stall samples (cuda,sass view: a source row, then its SASS rows)
    src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    per_line = defaultdict(lambda: [0, 0, ""])
    cur = None
    fname = ""
    cols = None
    for row in csv.reader(io.StringIO(src)):
        if not row:
            continue
        if row[0] in ("File Name", "File Path"):
            fname = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            cols = row
            continue
        if cols is None or len(row) < len(cols):
            continue
        if row[0]:
            cur = (fname, int(row[0]))
            per_line[cur][2] = row[1].strip()
            continue
        if cur is None:
            continue
        try:
            ie = int(row[cols.index("Instructions Executed")])
            sm = int(row[cols.index("# Samples")])
        except ValueError:
            continue
        per_line[cur][0] += ie
        per_line[cur][1] += sm
    tot_i = sum(v[0] for v in per_line.values()) or 1
    tot_s = sum(v[1] for v in per_line.values()) or 1
    print("== hottest source lines (warp instructions executed | stall samples)")
    for (f, ln), (ie, sm, text) in sorted(per_line.items(), key=lambda kv: -(kv[1][0] / tot_i + kv[1][1] / tot_s))[:28]:
        print("  %-16s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (f, ln, 100.0 * ie / tot_i, 100.0 * sm / tot_s, text[:90]))
    sass = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    lines = sass.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    tot_inst = 0
    tot_samp = 0
    ops = defaultdict(lambda: [0, 0])
    for r in rd:
        try:
            ie = int(r["Instructions Executed"])
            sm = int(r["# Samples"])
        except (KeyError, ValueError):
            continue
        tot_inst += ie
        tot_samp += sm
        op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
        ops[op.split(".")[0]][0] += ie
        ops[op.split(".")[0]][1] += sm
    print("== SASS opcode mix (warp instructions executed, stall samples)")
    for op, (ie, sm) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]:
        print("  %-12s %12d  %5.1f%%   samples %6d %5.1f%%" % (op, ie, 100.0 * ie / max(tot_inst, 1), sm, 100.0 * sm / max(tot_samp, 1)))
    print("  total        %12d" % tot_inst)


if __name__ == "__main__":
    main()
