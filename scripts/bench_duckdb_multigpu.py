#!/usr/bin/env python
"""read_fastq through DuckDB with one device pipeline per GPU (`gpus = n`): the same statements at 1 .. all GPUs of the box,
best of three runs each (the first run on a device pays its context / pool warm-up).
usage (multi-GPU box): python scripts/bench_duckdb_multigpu.py [--reads N] [--out gpurun_out/duckdb_mg.json]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SQLRUN = os.path.join(ROOT, "build", "rt", "sqlrun")
PRODUCT = os.path.join(ROOT, "exon_duckdb_b200", "duckdb_ext", "exon.duckdb_extension")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=16_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--out", default="gpurun_out/duckdb_mg.json")
    args = ap.parse_args()
    import torch
    from tools import synth
    n_dev = torch.cuda.device_count()
    path = os.path.join(args.dir, "exb_mg.fastq")
    synth.gen_host(synth.gen_params("illumina", args.reads, seed=20)).tofile(path)
    size = os.path.getsize(path)
    mq = "list_avg(quality_score_string_to_list(quality_scores)) > 30"
    queries = [
        ("COUNT(*) WHERE mean quality > 30", "SELECT COUNT(*) FROM read_fastq('%s', gpus=%d) WHERE " + mq),
        ("AVG(gc_content(sequence))", "SELECT AVG(gc_content(sequence)) FROM read_fastq('%s', gpus=%d)"),
        ("SUM(length(sequence))", "SELECT SUM(length(sequence)) FROM read_fastq('%s', gpus=%d)"),
        ("CREATE TABLE AS name, sequence (order preserved)", "CREATE OR REPLACE TABLE t AS SELECT name, sequence FROM read_fastq('%s', gpus=%d)"),
    ]
    rows = []
    gs = [g for g in (1, 2, 4, 8) if g <= n_dev]
    for name, q in queries:
        for g in gs:
            stmts = [q % (path, g)] * 3
            text = "LOAD '%s';\n" % PRODUCT + "\n".join(s + ";" for s in stmts) + "\n"
            out = subprocess.run([SQLRUN], input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1200)
            res = [json.loads(l) for l in out.stdout.decode().splitlines()][1:]
            ok = [r for r in res if r.get("ok")]
            if not ok:
                print("%-50s gpus=%d FAILED %s" % (name, g, res[:1]), flush=True)
                continue
            best = min(r["ms"] for r in ok)
            rows.append({"query": name, "gpus": g, "ms": best, "GB/s": size / (best * 1e-3) / 1e9, "all_ms": [r["ms"] for r in ok],
                         "result": ok[-1]["rows"][0][0] if ok[-1].get("rows") else None})
            print("%-50s gpus=%d  %8.1f ms  %6.2f GB/s   (runs: %s)" % (name, g, best, size / (best * 1e-3) / 1e9, " ".join("%.0f" % r["ms"] for r in ok)), flush=True)
    os.unlink(path)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"file_bytes": size, "devices": n_dev, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
