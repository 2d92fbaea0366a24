#!/usr/bin/env python
"""End-to-end throughput of the native reader (exb_reader_*: the calls the DuckDB extension's bind / init / scan
callbacks make, include/exon_b200.h section 1b) from a file on disk (tmpfs) to host-resident 2048-row batches.

usage (GPU box): python scripts/bench_reader.py [--reads N] [--out gpurun_out/reader.json]
Wall clock around open .. last batch .. close; GB/s = input file bytes / wall time.  The reference's equivalent is
new_reader + the Arrow stream pulled by ArrowScanParallelStateNext (arrow_table_function/module.cpp:216-294).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exon_duckdb_b200 import _lib
from tools import synth
from exon_duckdb_b200._lib import check, lib


def run(path, fmt, mask, filters, count_only, batch_rows=2048):
    h = C.c_void_p()
    t0 = time.perf_counter()
    check(lib().exb_reader_open(path.encode(), fmt.encode(), None, batch_rows, filters.encode() if filters else None, mask, C.byref(h)))
    rows = 0
    nbytes = 0
    if count_only:
        n = C.c_int64()
        check(lib().exb_reader_count(h, C.byref(n)))
        rows = n.value
    else:
        b = _lib.Batch()
        while True:
            check(lib().exb_reader_next(h, C.byref(b)))
            if b.n_rows == 0:
                break
            rows += b.n_rows
            for c in range(b.n_cols):
                if b.cols[c].offsets:
                    nbytes += b.cols[c].offsets[b.n_rows] - b.cols[c].offsets[0]
            lib().exb_batch_release(C.byref(b))
    lib().exb_reader_close(h)
    return time.perf_counter() - t0, rows, nbytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--contigs", type=int, default=1200)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--out", default="gpurun_out/reader.json")
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    from exon_duckdb_b200 import device as D

    rows = []
    fq = os.path.join(args.dir, "exb_bench.fastq")
    fa = os.path.join(args.dir, "exb_bench.fasta")
    synth.gen_host(synth.gen_params("illumina", args.reads, seed=20)).tofile(fq)
    synth.gen_host(synth.gen_params("fasta", args.contigs, seed=3, len_min=500000, len_max=500000, wrap=60)).tofile(fa)
    cases = [
        ("read_fastq COUNT(*)", fq, "fastq", 0xF, None, True),
        ("read_fastq COUNT(*) WHERE mean quality > 30", fq, "fastq", 0xF, "mean_quality(quality_scores) > 30", True),
        ("read_fastq all 4 columns", fq, "fastq", 0xF, None, False),
        ("read_fastq sequence only", fq, "fastq", 0x4, None, False),
        ("read_fastq name, sequence WHERE mean quality > 30", fq, "fastq", 0x5, "mean_quality(quality_scores) > 30", False),
        ("read_fasta COUNT(*)", fa, "fasta", 0x7, None, True),
        ("read_fasta id only", fa, "fasta", 0x1, None, False),
        ("read_fasta all 3 columns", fa, "fasta", 0x7, None, False),
    ]
    for name, path, fmt, mask, filt, cnt in cases:
        size = os.path.getsize(path)
        best = None
        try:
            for _ in range(args.repeat):
                dt, n, nb = run(path, fmt, mask, filt, cnt)
                best = dt if best is None or dt < best else best
        except Exception as e:  # a filter spelling the reader does not parse is reported, not fatal
            print("%-52s  FAILED: %s" % (name, e), flush=True)
            continue
        gbs = size / best / 1e9
        rows.append({"path": name, "file_bytes": size, "rows": n, "column_bytes": nb, "s_best": best, "GB/s": gbs})
        print("%-52s %8.1f ms  %7.2f GB/s  rows %d  column bytes %d" % (name, best * 1e3, gbs, n, nb), flush=True)
    os.unlink(fq)
    os.unlink(fa)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
