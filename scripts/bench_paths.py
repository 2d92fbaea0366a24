#!/usr/bin/env python
"""Device-resident throughput of every path of the scope table, one line each.

usage (GPU box): python scripts/bench_paths.py [--reads N] [--fasta-mb M] [--out gpurun_out/paths.json]
Times with CUDA events on the launching stream, inputs larger than L2, 3 warm-ups, best-of / mean of 10.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from exon_duckdb_b200 import _lib, device as D
from tools import synth


ITERS = None  # --iters overrides every call (profiling runs under ncu)
WARM = None


def timeit(fn, iters=10, warm=3):
    iters = ITERS or iters
    warm = warm if WARM is None else WARM
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--ont-reads", type=int, default=40_000)
    ap.add_argument("--contigs", type=int, default=2400)
    ap.add_argument("--contig-len", type=int, default=500_000)
    ap.add_argument("--out", default="gpurun_out/paths.json")
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--warm", type=int, default=-1)
    args = ap.parse_args()
    global ITERS, WARM
    ITERS = args.iters or None
    WARM = args.warm if args.warm >= 0 else None
    dev = torch.device("cuda:0")
    peak = 6548.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []

    def report(name, algo_bytes, ms_med, ms_best, note=""):
        gbs = algo_bytes / (ms_med * 1e-3) / 1e9
        rows.append({"path": name, "algorithmic_bytes": algo_bytes, "ms_median": ms_med, "ms_best": ms_best, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak,
                     "note": note})
        print("%-58s %9.3f ms  %8.1f GB/s  %5.1f%% of %.0f  %s" % (name, ms_med, gbs, 100 * gbs / peak, peak, note), flush=True)

    # ---------------- C2: Illumina FASTQ
    p = synth.gen_params("illumina", args.reads, seed=20)
    buf = synth.gen_device(p, dev)
    n = buf.numel()
    preds = [("mean_quality", ">", 30.0)]
    c = D.fastq_scan_filter(buf, preds)
    assert c.validate() == args.reads
    med, best = timeit(lambda: D.fastq_scan_filter(buf, preds, out=c))
    report("C2 fused scan+filter COUNT (exb_fastq_scan_filter)", n, med, best)
    rec_cap = args.reads + 1024
    for flags, nm in ((_lib.F_QUAL, "F_QUAL"), (_lib.F_SEQ | _lib.F_QUAL, "F_SEQ|F_QUAL"), (_lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, "F_LINES|F_SEQ|F_QUAL"),
                      (_lib.F_LINES, "F_LINES")):
        s = D.fastq_scan(buf, flags, rec_cap=rec_cap)
        assert s.validate() == args.reads
        out_b = args.reads * (8 * bool(flags & 2) + 8 * bool(flags & 4) + 16 * bool(flags & 1))
        med, best = timeit(lambda: D.fastq_scan(buf, flags, out=s))
        report("C2 general scan %s (exb_fastq_scan)" % nm, n + out_b, med, best)
        if flags == _lib.F_QUAL:
            agg = torch.zeros(8, dtype=torch.int64, device=dev)
            def scan_filter():
                D.fastq_scan(buf, flags, out=s)
                D.fastq_filter(s, rec_cap, preds, agg=agg, device_count=True)
            med, best = timeit(scan_filter)
            report("C2 general scan F_QUAL + exb_fastq_filter COUNT", n, med, best)
    del s
    tab = D.fastq_table(buf, columns=["name", "sequence"], preds=preds)
    out_b = tab["name"].data.numel() + tab["sequence"].data.numel() + 16 * tab["__n_rows__"]
    med, best = timeit(lambda: D.fastq_table(buf, columns=["name", "sequence"], preds=preds), iters=5)
    report("C2 filter projecting name+sequence (fastq_table)", n + out_b, med, best, "includes host syncs for sizes")
    tab = D.fastq_table(buf)
    out_b = sum(tab[k].data.numel() for k in D.FASTQ_COLUMNS) + 32 * tab["__n_rows__"]
    med, best = timeit(lambda: D.fastq_table(buf), iters=5)
    report("C2 full 4-column materialisation (fastq_table)", n + out_b, med, best, "includes host syncs for sizes")
    seq = tab["sequence"]
    qual = tab["quality_scores"]
    med, best = timeit(lambda: D.gc_content(seq))
    report("gc_content(sequence) over a column (exb_gc_content)", seq.data.numel() + 12 * len(seq), med, best)
    med, best = timeit(lambda: D.reverse_complement(seq))
    report("reverse_complement(sequence) (exb_seq_map)", 2 * seq.data.numel(), med, best, "includes 1 host sync for the error flag")
    med, best = timeit(lambda: D.quality_score_string_to_list(qual))
    report("quality_score_string_to_list (exb_quality_decode)", 5 * qual.data.numel(), med, best)
    del tab, seq, qual, buf, c
    torch.cuda.empty_cache()

    # ---------------- C4: ONT FASTQ, reverse_complement projection
    p = synth.gen_params("ont", args.ont_reads, seed=4, len_min=10000, len_max=50000)
    buf = synth.gen_device(p, dev)
    n = buf.numel()
    s = D.fastq_scan(buf, _lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, rec_cap=args.ont_reads + 1024)
    assert s.validate() == args.ont_reads
    med, best = timeit(lambda: D.fastq_scan(buf, _lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, out=s))
    report("C4 ONT general scan F_LINES|F_SEQ|F_QUAL", n, med, best)
    def c4():
        t = D.fastq_table(buf, columns=["sequence"])
        return D.reverse_complement(t["sequence"])
    rc = c4()
    med, best = timeit(c4, iters=5)
    # algorithmic bytes: input once + output strings once, however many passes the implementation makes
    report("C4 read_fastq -> reverse_complement(sequence)", n + rc.data.numel(), med, best, "scan + gather + LUT map kernel, host syncs included")
    def c4f():
        return D.fastq_table(buf, columns=["sequence"], seq_map="reverse_complement")["sequence"]
    rcf = c4f()
    assert torch.equal(rcf.data, rc.data)
    med, best = timeit(c4f, iters=5)
    report("C4 the same, LUT fused into the gather (exb_fastq_gather_map)", n + rcf.data.numel(), med, best, "scan + one gather; host syncs included")
    del rcf
    del rc, s, buf
    torch.cuda.empty_cache()

    # ---------------- C3: wrapped FASTA, gc_content per contig
    p = synth.gen_params("fasta", args.contigs, seed=3, len_min=args.contig_len, len_max=args.contig_len, wrap=60)
    buf = synth.gen_device(p, dev)
    n = buf.numel()
    fs = D.fasta_scan(buf, compact=False)
    assert int(fs.result.n_records) == args.contigs
    def c3():
        D.fasta_scan(buf, compact=False, out=fs)
        return D.gc_from_prefix(fs.seq_off, fs.gc_prefix, args.contigs)
    med, best = timeit(c3)
    report("C3 read_fasta + gc_content per contig (no sequence column)", n, med, best)
    fs2 = D.fasta_scan(buf, compact=True)
    seq_bytes = int(fs2.result.seq_bytes)
    med, best = timeit(lambda: D.fasta_scan(buf, compact=True, out=fs2))
    report("C3 read_fasta with the sequence column compacted", n + seq_bytes, med, best)
    with open(args.out, "w") as f:
        json.dump({"peak_gbs": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
