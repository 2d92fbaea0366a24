"""Device-resident throughput of every path of the scope table, one row each: what bench.py reports
under `paths` and scripts/bench_paths.py prints.  CUDA events on the launching stream, inputs larger than L2, warm-ups,
median of the timed iterations.  Algorithmic bytes follow SURVEY 8(d): input once + output once."""
import json
import os

import torch

from exon_duckdb_b200 import _lib, device as D
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def committed_traffic():
    """DRAM bytes per launch from the committed ncu captures (profiles/traffic_paths.json), keyed by path name, with the
    input size the capture was taken on: {"name": {"dram_bytes": .., "input_bytes": ..}}."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_paths.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def timeit(fn, iters=10, warm=3):
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


class Report:
    def __init__(self, peak=None, verbose=True):
        self.peak = peak or hbm_peak()
        self.rows = []
        self.verbose = verbose
        self.traffic = committed_traffic()

    def add(self, name, algo_bytes, ms_med, ms_best, input_bytes, note=""):
        gbs = algo_bytes / (ms_med * 1e-3) / 1e9
        tr = self.traffic.get(name)
        traffic = None
        if tr and tr.get("dram_bytes") and tr.get("input_bytes"):
            traffic = tr["dram_bytes"] * (input_bytes / tr["input_bytes"])  # streaming kernels: traffic scales with the input
        self.rows.append({"path": name, "algorithmic_bytes": algo_bytes, "input_bytes": input_bytes, "ms_median": ms_med, "ms_best": ms_best,
                          "GB/s": gbs, "frac": gbs / self.peak, "traffic": traffic, "note": note})
        if self.verbose:
            print("%-62s %9.3f ms  %8.1f GB/s  %5.1f%% of %.0f  %s" % (name, ms_med, gbs, 100 * gbs / self.peak, self.peak, note), flush=True)


def c2_paths(rep, buf, reads, iters=10, full=True):
    """Illumina FASTQ already resident in `buf`."""
    n = buf.numel()
    preds = [("mean_quality", ">", 30.0)]
    rec_cap = reads + 1024
    if full and not os.environ.get("EXB_PATHS_SPLIT_ONLY"):
        c = D.fastq_scan_filter(buf, preds)
        assert c.validate() == reads
        med, best = timeit(lambda: D.fastq_scan_filter(buf, preds, out=c), iters)
        rep.add("C2 fused scan+filter COUNT (exb_fastq_scan_filter)", n, med, best, n)
    split_only = bool(os.environ.get("EXB_PATHS_SPLIT_ONLY"))
    variants = () if split_only else ((_lib.F_QUAL, "F_QUAL"), (_lib.F_SEQ | _lib.F_QUAL, "F_SEQ|F_QUAL"), (_lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, "F_LINES|F_SEQ|F_QUAL"),
                (_lib.F_LINES, "F_LINES")) if full else ((_lib.F_SEQ | _lib.F_QUAL, "F_SEQ|F_QUAL"),)
    for flags, nm in variants:
        s = D.fastq_scan(buf, flags, rec_cap=rec_cap)
        assert s.validate() == reads
        out_b = reads * (8 * bool(flags & 2) + 8 * bool(flags & 4) + 16 * bool(flags & 1))
        med, best = timeit(lambda: D.fastq_scan(buf, flags, out=s), iters)
        rep.add("C2 general scan %s (exb_fastq_scan)" % nm, n + out_b, med, best, n)
        if flags == _lib.F_QUAL:
            agg = torch.zeros(8, dtype=torch.int64, device=buf.device)

            def scan_filter():
                D.fastq_scan(buf, flags, out=s)
                D.fastq_filter(s, rec_cap, preds, agg=agg, device_count=True)
            med, best = timeit(scan_filter, iters)
            rep.add("C2 general scan F_QUAL + exb_fastq_filter COUNT", n, med, best, n)
    if not split_only:
        del s
        tab = D.fastq_table(buf, columns=["name", "sequence"], preds=preds)
        out_b = tab["name"].data.numel() + tab["sequence"].data.numel() + 16 * tab["__n_rows__"]
        med, best = timeit(lambda: D.fastq_table(buf, columns=["name", "sequence"], preds=preds), iters=5)
        rep.add("C2 filter projecting name+sequence (fastq_table)", n + out_b, med, best, n, "includes host syncs for sizes")
    tab = D.fastq_table(buf)
    out_b = sum(tab[k].data.numel() for k in D.FASTQ_COLUMNS) + 32 * tab["__n_rows__"]
    med, best = timeit(lambda: D.fastq_table(buf), iters=5)
    rep.add("C2 full 4-column materialisation (fastq_table)", n + out_b, med, best, n, "includes host syncs for sizes")
    # the same work with the outputs allocated up front and nothing read back in between: what the device itself takes
    # (the reader allocates its chunk buffers once, too).  exb_fastq_scan(F_LINES) + exb_fastq_split, all four columns.
    import ctypes as C
    from exon_duckdb_b200._lib import check, lib
    dev = buf.device
    s = D.fastq_scan(buf, _lib.F_LINES, rec_cap=rec_cap)
    assert s.validate() == reads
    offs = torch.empty((4, reads + 1), dtype=torch.int64, device=dev)
    valid = torch.empty(reads, dtype=torch.uint8, device=dev)
    scratch = torch.empty(lib().exb_fastq_split_scratch_bytes(reads), dtype=torch.uint8, device=dev)
    caps = [int(tab[k].data.numel()) + 64 for k in D.FASTQ_COLUMNS]
    data = [torch.empty(c + 16, dtype=torch.uint8, device=dev) for c in caps]
    outs = (C.c_void_p * 4)(*[D._ptr(d) for d in data])
    capv = (C.c_int64 * 4)(*caps)

    def direct():
        D.fastq_scan(buf, _lib.F_LINES, out=s)
        check(lib().exb_fastq_split(D._ptr(buf), 0, n, D._ptr(s.line_end), 1 if s.wide else 0, reads, 0xF, D._ptr(offs), D._ptr(valid), outs, capv,
                                    D._ptr(scratch), D._ptr(s.ws), -1, None, D._stream()))
    direct()
    torch.cuda.synchronize()
    assert offs[:, reads].cpu().tolist() == [int(tab[k].data.numel()) for k in D.FASTQ_COLUMNS]
    assert all(torch.equal(data[c][:caps[c] - 64], tab[k].data) for c, k in enumerate(D.FASTQ_COLUMNS))
    med, best = timeit(direct, iters)
    rep.add("C2 full 4-column materialisation, device only (exb_fastq_scan + exb_fastq_split)", n + out_b, med, best, n, "outputs preallocated, no host round trip")
    med, best = timeit(lambda: check(lib().exb_fastq_split(D._ptr(buf), 0, n, D._ptr(s.line_end), 1 if s.wide else 0, reads, 0xF, D._ptr(offs), D._ptr(valid),
                                                           outs, capv, D._ptr(scratch), D._ptr(s.ws), -1, None, D._stream())), iters)
    rep.add("C2 column split alone (exb_fastq_split: fields + offsets + 4 columns)", out_b + (out_b - 32 * reads) + 16 * reads, med, best, n,
            "reads 4 B per line + the record bytes, writes offsets + columns")
    del data, offs, valid, scratch, s
    if full and not split_only:
        seq = tab["sequence"]
        qual = tab["quality_scores"]
        med, best = timeit(lambda: D.gc_content(seq), iters)
        rep.add("gc_content(sequence) over a column (exb_gc_content)", seq.data.numel() + 12 * len(seq), med, best, seq.data.numel())
        med, best = timeit(lambda: D.reverse_complement(seq), iters)
        rep.add("reverse_complement(sequence) (exb_seq_map)", 2 * seq.data.numel(), med, best, seq.data.numel(), "includes 1 host sync for the error flag")
        med, best = timeit(lambda: D.quality_score_string_to_list(qual), iters)
        rep.add("quality_score_string_to_list (exb_quality_decode)", 5 * qual.data.numel(), med, best, qual.data.numel())
    if not split_only:
        # the writer (COPY ... TO (FORMAT 'fastq')): the four columns back into a file image -- which must be the input again
        cols = [tab[k] for k in D.FASTQ_COLUMNS]
        img, _ = D.fastq_format(*cols)
        assert img.numel() == n and torch.equal(img, buf[:n])
        col_b = sum(c.data.numel() for c in cols) + 32 * len(cols[0])
        del img
        med, best = timeit(lambda: D.fastq_format(*cols), iters=5)
        rep.add("C2 writer: 4 columns -> FASTQ file image (exb_fastq_format)", col_b + n, med, best, n, "includes the scratch / image allocations and 1 host sync")
    del tab


def c4_paths(rep, dev, ont_reads, iters=5):
    p = synth.gen_params("ont", ont_reads, seed=4, len_min=10000, len_max=50000)
    buf = synth.gen_device(p, dev)
    n = buf.numel()
    s = D.fastq_scan(buf, _lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, rec_cap=ont_reads + 1024)
    assert s.validate() == ont_reads
    med, best = timeit(lambda: D.fastq_scan(buf, _lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, out=s), iters)
    rep.add("C4 ONT general scan F_LINES|F_SEQ|F_QUAL", n, med, best, n)

    def c4f():
        return D.fastq_table(buf, columns=["sequence"], seq_map="reverse_complement")["sequence"]
    rcf = c4f()
    med, best = timeit(c4f, iters)
    # algorithmic bytes: input once + output strings once, however many passes the implementation makes
    rep.add("C4 read_fastq -> reverse_complement(sequence), LUT fused into the gather", n + rcf.data.numel(), med, best, n, "scan + one gather; host syncs included")
    del rcf, s, buf
    torch.cuda.empty_cache()


def c3_paths(rep, dev, contigs, contig_len, iters=10):
    p = synth.gen_params("fasta", contigs, seed=3, len_min=contig_len, len_max=contig_len, wrap=60)
    buf = synth.gen_device(p, dev)
    n = buf.numel()
    fs = D.fasta_scan(buf, compact=False)
    assert int(fs.result.n_records) == contigs

    def c3():
        D.fasta_scan(buf, compact=False, out=fs)
        return D.gc_from_prefix(fs.seq_off, fs.gc_prefix, contigs)
    med, best = timeit(c3, iters)
    rep.add("C3 read_fasta + gc_content per contig (no sequence column)", n, med, best, n)
    fs2 = D.fasta_scan(buf, compact=True)
    seq_bytes = int(fs2.result.seq_bytes)
    med, best = timeit(lambda: D.fasta_scan(buf, compact=True, out=fs2), iters)
    rep.add("C3 read_fasta with the sequence column compacted", n + seq_bytes, med, best, n)
    # the writer (COPY ... TO (FORMAT 'fasta')): id / description / sequence columns -> file image re-wrapped at 60
    tab = D.fasta_table(buf)
    cols = [tab[k] for k in D.FASTA_COLUMNS]
    img, _ = D.fasta_format(*cols, line_width=60)
    assert img.numel() == n and torch.equal(img, buf[:n])
    del img
    med, best = timeit(lambda: D.fasta_format(*cols, line_width=60), iters=5)
    rep.add("C3 writer: 3 columns -> FASTA file image wrapped at 60 (exb_fasta_format)", seq_bytes + n, med, best, n, "includes the scratch / image allocations and 1 host sync")
    del tab, cols, fs, fs2, buf
    torch.cuda.empty_cache()


def bgzf_image(text, level=6, block=65280, threads=16):
    """BGZF file image of `text` (what bgzip writes: members of 65280 text bytes, zlib level 6), compressed on `threads`
    host threads (zlib releases the GIL).  Tooling: benchmark and test input."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    mv = memoryview(text)

    def member(lo):
        piece = mv[lo:lo + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        z = c.compress(piece) + c.flush()
        hdr = b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff\x06\x00BC\x02\x00" + (len(z) + 25).to_bytes(2, "little")
        return hdr + z + zlib.crc32(piece).to_bytes(4, "little") + len(piece).to_bytes(4, "little")

    with ThreadPoolExecutor(threads) as ex:
        return b"".join(ex.map(member, range(0, len(text), block)))


def bgzf_paths(rep, dev, reads=400_000, copies=8, iters=5):
    """bgzip'ed FASTQ inflated on the device (exb_bgzf_inflate, CRC-32 checked), `copies` x `reads`
    Illumina reads so that text + compressed bytes exceed L2."""
    import ctypes as C
    import numpy as np

    text = synth.gen_host(synth.gen_params("illumina", reads, seed=20)).tobytes()
    img = bgzf_image(text) * copies
    tab, n, nxt, outb = D.bgzf_index(img)
    assert nxt == len(img) and outb == len(text) * copies
    d_in = D.to_device(img, dev)
    d_tab = torch.from_numpy(np.frombuffer(bytes(tab), dtype=np.uint8)[: n * 32].copy()).to(dev)
    d_out = D.alloc_input(outb, dev)
    d_state = torch.empty(16, dtype=torch.uint8, device=dev)
    L = _lib.lib()

    def run(crc):
        _lib.check(L.exb_bgzf_inflate(D._ptr(d_in), D._ptr(d_tab), n, D._ptr(d_out), D._ptr(d_state), crc, D._stream()))

    for crc, nm in ((1, "CRC-32 checked"), (0, "no CRC")):
        med, best = timeit(lambda: run(crc), iters=iters)
        _lib.check(L.exb_bgzf_finish(D._ptr(d_state), None, D._stream()))
        rep.add("BGZF FASTQ inflated on the device (exb_bgzf_inflate, %s)" % nm, len(img) + outb, med, best, len(img),
                "%d members, %.2f GB text from %.2f GB (ratio %.2f); text GB/s = %.1f" % (n, outb / 1e9, len(img) / 1e9, outb / len(img), outb / med / 1e6))
    got = d_out[: len(text)].cpu().numpy().tobytes()
    assert got == text
    del d_in, d_out, d_tab
    torch.cuda.empty_cache()
