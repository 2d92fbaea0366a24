#!/usr/bin/env python
"""End-to-end throughput of the native reader (exb_reader_*: the calls the DuckDB extension's bind / init / scan
callbacks make, include/exon_b200.h) from a file on disk (tmpfs) to host-resident 2048-row batches in DuckDB's
vector layout.

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
from exon_duckdb_b200._lib import check, lib
from tools import synth


def run(path, fmt, mask, filters, count_only, computed=(), batch_rows=2048, copy_io=False):
    h = C.c_void_p()
    t0 = time.perf_counter()
    o = _lib.reader_options(column_mask=mask, flags=_lib.RD_STRING_T | _lib.RD_NO_OFFSETS | (_lib.RD_COPY_IO if copy_io else 0), computed=computed)
    check(lib().exb_reader_open2(path.encode(), fmt.encode(), None, batch_rows, filters.encode() if filters else None, C.byref(o), C.byref(h)))
    rows = 0
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
            lib().exb_batch_release(C.byref(b))
    lib().exb_reader_close(h)
    return time.perf_counter() - t0, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--contigs", type=int, default=2400)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--out", default="gpurun_out/reader.json")
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--cases", default="plain,bgzf")
    args = ap.parse_args()

    rows = []
    fq = os.path.join(args.dir, "exb_bench.fastq")
    fa = os.path.join(args.dir, "exb_bench.fasta")
    if not os.path.exists(fq):
        synth.gen_host(synth.gen_params("illumina", args.reads, seed=20)).tofile(fq)
    if not os.path.exists(fa):
        synth.gen_host(synth.gen_params("fasta", args.contigs, seed=3, len_min=500000, len_max=500000, wrap=60)).tofile(fa)
    cases = [
        ("read_fastq COUNT(*)", fq, "fastq", 0, None, True, ()),
        ("read_fastq COUNT(*) WHERE mean quality > 30", fq, "fastq", 0, "mean_quality(quality_scores)>30", True, ()),
        ("read_fastq all 4 columns", fq, "fastq", 0xF, None, False, ()),
        ("read_fastq sequence only", fq, "fastq", 0x4, None, False, ()),
        ("read_fastq name, sequence WHERE mean quality > 30", fq, "fastq", 0x5, "mean_quality(quality_scores)>30", False, ()),
        ("read_fastq gc_content(sequence) [computed, no column read back]", fq, "fastq", 0, None, False, ((_lib.C_GC_CONTENT, 0),)),
        ("read_fastq reverse_complement(sequence) [computed]", fq, "fastq", 0, None, False, ((_lib.C_SEQ_MAP, 0),)),
        ("read_fastq quality_score_string_to_list(quality_scores) [computed]", fq, "fastq", 0, None, False, ((_lib.C_QUALITY_LIST, 0),)),
        ("read_fasta COUNT(*)", fa, "fasta", 0, None, True, ()),
        ("read_fasta id, gc_content(sequence) [computed]", fa, "fasta", 0x1, None, False, ((_lib.C_GC_CONTENT, 0),)),
        ("read_fasta all 3 columns", fa, "fasta", 0x7, None, False, ()),
    ]
    for name, path, fmt, mask, filt, cnt, comp in (cases if "plain" in args.cases else []):
        size = os.path.getsize(path)
        # two I/O paths: "copy" = what the FIRST scan of a file does (page cache -> pinned blocks -> DMA); "registered" = every
        # later scan of a memory-backed file (its page cache was pinned in place by cudaHostRegister after the first scan)
        for io in ("copy", "registered"):
            if io == "registered":
                run(path, fmt, mask, filt, cnt, comp)  # a complete scan without the flag triggers the registration
                t0 = time.time()
                while lib().exb_file_cache_state(path.encode()) == 1 and time.time() - t0 < 60:
                    time.sleep(0.01)
                if lib().exb_file_cache_state(path.encode()) != 2:
                    print("%-68s registered page cache not available here" % name, flush=True)
                    continue
            times = []
            for _ in range(args.repeat):
                dt, n = run(path, fmt, mask, filt, cnt, comp, copy_io=(io == "copy"))
                times.append(dt)
            best, med = min(times), sorted(times)[len(times) // 2]
            rows.append({"case": name, "io": io, "bytes": size, "rows": n, "best_ms": best * 1e3, "median_ms": med * 1e3, "best_gbs": size / 1e9 / best,
                         "median_gbs": size / 1e9 / med, "all_ms": [t * 1e3 for t in times]})
            print("%-68s %-10s best %7.1f ms %6.2f GB/s | median %7.1f ms %6.2f GB/s  rows %d" % (name, io, best * 1e3, size / 1e9 / best, med * 1e3, size / 1e9 / med, n), flush=True)
    # ---- bgzip'ed FASTQ: the members cross PCIe compressed and are inflated on the device; the streaming
    # zlib decoder (EXON_B200_BGZF=0: what every other gzip file takes, and what the reference does) beside it
    from tools import paths as P
    fqz = os.path.join(args.dir, "exb_bench_bgzf.fastq.gz")
    text_bytes = os.path.getsize(fq)
    if not os.path.exists(fqz):
        with open(fq, "rb") as f:
            img = P.bgzf_image(f.read())
        with open(fqz, "wb") as f:
            f.write(img)
        del img
    zbytes = os.path.getsize(fqz)
    zcases = [
        ("read_fastq(bgzf) COUNT(*) WHERE mean quality > 30", 0, "mean_quality(quality_scores)>30", True, ()),
        ("read_fastq(bgzf) all 4 columns", 0xF, None, False, ()),
        ("read_fastq(bgzf) gc_content(sequence) [computed]", 0, None, False, ((_lib.C_GC_CONTENT, 0),)),
    ]
    for name, mask, filt, cnt, comp in (zcases if "bgzf" in args.cases else []):
        for mode, reps in (("device inflate", args.repeat), ("zlib stream", 1)):
            if mode == "zlib stream":
                if not cnt:
                    continue
                os.environ["EXON_B200_BGZF"] = "0"
            else:
                os.environ.pop("EXON_B200_BGZF", None)
            times = []
            for _ in range(reps + (1 if mode == "device inflate" else 0)):
                dt, n = run(fqz, "fastq", mask, filt, cnt, comp)
                times.append(dt)
            times = times[1:] if len(times) > 1 else times
            best = min(times)
            rows.append({"case": name, "io": mode, "bytes": zbytes, "text_bytes": text_bytes, "rows": n, "best_ms": best * 1e3,
                         "best_gbs": zbytes / 1e9 / best, "best_text_gbs": text_bytes / 1e9 / best, "all_ms": [t * 1e3 for t in times]})
            print("%-68s %-14s best %8.1f ms %6.2f GB/s of file bytes = %6.2f GB/s of text  rows %d" %
                  (name, mode, best * 1e3, zbytes / 1e9 / best, text_bytes / 1e9 / best, n), flush=True)
    os.environ.pop("EXON_B200_BGZF", None)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
