#!/usr/bin/env python
"""Per-source-line warp-instruction counts of an .ncu-rep, normalised per launched warp.

usage: python profiles/ncu_lines.py rep.ncu-rep [min_inst_per_warp]
"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0  # normalise per this many work units (e.g. tiles) instead of per warp
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
warps = int(d["launch__grid_size"].replace(",", "")) * int(d["launch__block_size"].replace(",", "")) // 32
if units:
    warps = units
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
per = defaultdict(lambda: [0, 0, ""])
cur = None; fname = ""; cols = None
for row in csv.reader(io.StringIO(src)):
    if not row: continue
    if row[0] in ("File Name", "File Path"): fname = row[1].split("/")[-1]; continue
    if row[0] == "Line No": cols = row; continue
    if cols is None or len(row) < len(cols): continue
    if row[0]:
        cur = (fname, int(row[0])); per[cur][2] = row[1].strip(); continue
    if cur is None: continue
    try:
        ie = int(row[cols.index("Instructions Executed")]); sm = int(row[cols.index("# Samples")])
    except ValueError: continue
    per[cur][0] += ie; per[cur][1] += sm
ti = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values()) or 1
print("warps %d  inst/warp %.1f  samples %d" % (warps, ti / warps, ts))
for (f, ln), (ie, sm, text) in sorted(per.items()):
    if ie / warps >= thr or sm / ts > 0.01:
        print("%-16s %4d %8.1f i/w %5.1f%% smp  %s" % (f, ln, ie / warps, 100.0 * sm / ts, text[:100]))
