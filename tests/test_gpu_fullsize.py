"""GPU parity at BASELINE.json's FULL sizes, through properties that do not need the oracle to read gigabytes:

  * two independent CUDA paths agree (fused COUNT vs general scan + filter kernel; scan prefixes vs scalar kernels),
  * linearity / sharding invariance (halves and byte-range shards cut inside records sum to the whole),
  * conservation (bytes in = bytes out of the compaction; every read of the generator is found, with its length),
  * involution (the reference's reverse_complement table applied twice is the identity),
  * and the oracle itself on bounded samples cut from the MIDDLE of the same generated file.

Sizes follow SURVEY 8(d): C2 = 20 M x 150 bp Illumina reads (7 GB), C3 = 6 000 x 500 kb contigs wrapped at 60 (3 GB),
C4 = 200 000 ONT reads of 10-50 kb (12 GB).  Set EXB_FULLSIZE_SCALE (default 1.0) to shrink them on a small GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu

SCALE = float(os.environ.get("EXB_FULLSIZE_SCALE", "1.0"))
PREDS = [("mean_quality", ">", 30.0)]


def _aligned(D, view):
    """The C ABI wants 16-byte aligned buffers with slack behind them: copy a misaligned view."""
    if view.data_ptr() % 16 == 0:
        return view
    out = D.alloc_input(view.numel(), view.device)
    out.copy_(view)
    return out


def _free(torch):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def test_c2_illumina_20m_reads(cuda_device):
    import torch
    from exon_duckdb_b200 import _lib, device as D, dist
    from oracle import oracle as O

    reads = max(40_000, int(20_000_000 * SCALE))
    p = synth.gen_params("illumina", reads, seed=20)
    buf = synth.gen_device(p, cuda_device)
    n = buf.numel()

    # (1) the fused COUNT kernel and the general scan + filter kernel are different code: same answer
    c = D.fastq_scan_filter(buf, PREDS)
    assert c.validate() == reads
    fused = c.agg.cpu().tolist()
    scan = D.fastq_scan_sync(buf, _lib.F_SEQ | _lib.F_QUAL, rec_cap=reads + 1024)
    assert scan.validate() == reads
    agg, _ = D.fastq_filter(scan, reads, PREDS)
    gen = agg.cpu().tolist()
    assert (fused[0], fused[3], fused[4]) == (gen[0], gen[3], gen[4])
    assert 0.4 * reads < fused[0] < 0.8 * reads
    # (2) conservation: every read is 150 bases with 150 qualities, G/C about half
    every, _ = D.fastq_filter(scan, reads, [])
    e = every.cpu().tolist()
    assert e[0] == reads and e[1] == 150 * reads and e[4] == 150 * reads and 0.49 < e[2] / e[1] < 0.51
    assert int(scan.seq_len[:reads].min()) == 150 and int(scan.seq_len[:reads].max()) == 150
    del scan, agg, every
    _free(torch)

    # (3) linearity: two halves cut at a record boundary sum to the whole
    half = synth.gen_params("illumina", reads // 2, seed=20)
    cut = int(synth.gen_size(half))
    parts = [D.fastq_scan_filter(buf[:cut], PREDS), D.fastq_scan_filter(_aligned(D, buf[cut:]), PREDS)]
    assert parts[0].validate() + parts[1].validate() == reads
    s = [a + b for a, b in zip(parts[0].agg.cpu().tolist(), parts[1].agg.cpu().tolist())]
    assert (s[0], s[3], s[4]) == (fused[0], fused[3], fused[4])
    del parts
    _free(torch)

    # (4) sharding invariance: three byte-range shards whose edges fall inside records, phase resolved from the
    # exchanged result blocks (exon_duckdb_b200/dist.py), sum to the whole
    G = 3
    bounds = [dist.byte_range(n, k, G)[0] & ~15 for k in range(G)] + [n]
    shards = []
    for k in range(G):
        lo, hi = bounds[k], bounds[k + 1]
        begin = 0 if k == 0 else dist.HALO
        shards.append(dist.Shard(buf[lo - begin:hi], lo, hi, begin, k == G - 1))  # lo is a multiple of 16: views stay aligned
    ranges = [[s_.lo, s_.hi, s_.begin] for s_ in shards]
    jobs = [dist.ShardedFastqCount(s_, PREDS, None, ranges=ranges) for s_ in shards]
    blocks = torch.cat([j.scan().clone() for j in jobs])
    total = torch.zeros(8, dtype=torch.int64, device=cuda_device)
    for k, j in enumerate(jobs):
        total += j.resolve(blocks, k)
    t = dist.check_count(total)
    assert (t[0], t[3], t[4]) == (fused[0], fused[3], fused[4])
    del jobs, shards, blocks
    _free(torch)

    # (5) the oracle on a bounded sample from the MIDDLE of the file: records [reads/2, reads/2 + 100k)
    k = min(100_000, reads // 4)
    mid = synth.gen_params("illumina", k, seed=20, first_record=reads // 2)
    text = synth.gen_host(mid).tobytes()
    assert buf[cut:cut + len(text)].cpu().numpy().tobytes() == text  # the device generator and the host generator agree
    sample = D.fastq_scan_filter(_aligned(D, buf[cut:cut + len(text)]), PREDS)
    assert sample.validate() == k
    want = O.fastq_count_mean_quality(text, ">", 30.0)
    assert sample.agg.cpu().tolist()[0] == want[0]


def test_c3_genome_3gbp(cuda_device):
    import torch
    from exon_duckdb_b200 import _lib, device as D
    from oracle import oracle as O

    contigs = max(12, int(6000 * SCALE))
    L = 500_000
    p = synth.gen_params("fasta", contigs, seed=3, len_min=L, len_max=L, wrap=60)
    buf = synth.gen_device(p, cuda_device)
    s = D.fasta_scan_sync(buf, rec_cap=contigs + 16, compact=True)
    r = s.result
    assert int(r.n_records) == contigs and int(r.seq_bytes) == contigs * L
    off = s.seq_off[:contigs + 1]
    assert torch.equal(off, torch.arange(contigs + 1, device=cuda_device, dtype=torch.int64) * L)
    # conservation: the compacted column has no line terminator left, and as many bytes as the input minus headers / LFs
    seq = s.seq[:contigs * L]
    assert int((seq == 10).sum()) == 0 and int((seq == 13).sum()) == 0
    # two paths to gc_content per contig: the scan's prefix counts vs the scalar kernel over the compacted column
    gc_scan = D.gc_from_prefix(s.seq_off, s.gc_prefix, contigs)
    gc_col = D.gc_content(D.Column(off, seq))
    assert torch.equal(gc_scan, gc_col)
    assert 0.30 < float(gc_scan.min()) and float(gc_scan.max()) < 0.70  # the generator draws U[0.35, 0.65] per contig
    assert int(s.gc_prefix[contigs]) == int(((seq == ord("G")) | (seq == ord("C"))).sum())
    # the oracle on contig #contigs/2, cut out of the middle of the file
    k = contigs // 2
    hs, he = int(s.hdr_start[k]), int(s.hdr_start[k + 1])
    text = buf[hs:he].cpu().numpy().tobytes()
    ref = O.parse_fasta(text)
    assert ref.n == 1 and ref.strings("sequence")[0] == seq[k * L:(k + 1) * L].cpu().numpy().tobytes()
    assert np.float32(O.gc_content(ref.strings("sequence")[0])) == np.float32(gc_scan[k].item())


def test_c4_ont_200k_reads(cuda_device):
    import torch
    from exon_duckdb_b200 import _lib, device as D
    from oracle import oracle as O

    reads = max(400, int(200_000 * SCALE))
    p = synth.gen_params("ont", reads, seed=4, len_min=10_000, len_max=50_000)
    buf = synth.gen_device(p, cuda_device)
    scan = D.fastq_scan_sync(buf, _lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL, rec_cap=reads + 1024)
    assert scan.validate() == reads
    sl, ql = scan.seq_len[:reads], scan.qual_len[:reads]
    assert torch.equal(sl, ql) and int(sl.min()) >= 10_000 and int(sl.max()) <= 50_000
    total = int(sl.sum())
    del scan
    _free(torch)
    # the projection both ways: gather + LUT kernel, and the LUT fused into the gather
    col = D.fastq_table(buf, columns=["sequence"])["sequence"]
    assert col.data.numel() == total
    rc = D.reverse_complement(col)
    fused = D.fastq_table(buf, columns=["sequence"], seq_map="reverse_complement")["sequence"]
    assert torch.equal(rc.data, fused.data) and torch.equal(rc.offsets, fused.offsets)
    del fused
    # involution: A->C->A, T->G->T, C->A->C, G->T->G
    back = D.reverse_complement(rc)
    assert torch.equal(back.data, col.data)
    # the oracle on one read from the middle
    k = reads // 2
    a, b = int(col.offsets[k]), int(col.offsets[k + 1])
    s = col.data[a:b].cpu().numpy().tobytes()
    assert rc.data[a:b].cpu().numpy().tobytes() == O.reverse_complement(s)


def test_c5_sharded_count_and_gc(cuda_device):
    """C5 (byte-range shards, COUNT + gc_content): eight shards cut at arbitrary bytes agree on record-aligned bounds
    from their exchanged states (dist.fastq_record_bounds: newline counts give the exact phase), every shard then
    scans its own records IN PLACE (unaligned `begin`), and COUNT / SUM(len) / SUM(#GC) add up to the single-shot
    scan exactly, AVG(gc_content) to within float summation order."""
    import torch
    from exon_duckdb_b200 import _lib, device as D, dist

    reads = max(80_000, int(6_000_000 * SCALE))
    G = 8
    buf = synth.gen_device(synth.gen_params("illumina", reads, seed=20), cuda_device)
    n = buf.numel()
    whole = D.fastq_scan_sync(buf, _lib.F_SEQ, rec_cap=reads + 1024)
    assert whole.validate() == reads
    agg, _ = D.fastq_filter(whole, reads, [])
    want = agg.cpu().tolist()[:3]
    want_gc = float(D.gc_from_counts(whole.seq_len, whole.gc, reads).double().sum())
    del whole, agg
    _free(torch)

    los = [dist.byte_range(n, k, G)[0] & ~15 for k in range(G)] + [n]
    states = []
    for k in range(G):
        lo, hi = los[k], los[k + 1]
        begin = 0 if k == 0 else dist.HALO
        states.append(dist.fastq_shard_state(dist.Shard(buf[lo - begin:hi], lo, hi, begin, k == G - 1)))
    bounds = dist.fastq_record_bounds(states, n)
    assert bounds[0] == 0 and bounds[-1] == n and bounds == sorted(bounds)
    assert all(los[k] <= bounds[k] < los[k] + 400 for k in range(G))  # the next record starts within one record's length
    got = [0, 0, 0]
    got_gc = 0.0
    for k in range(G):
        a, b = bounds[k], bounds[k + 1]
        if a == b:
            continue
        s = D.fastq_scan_sync(buf, _lib.F_SEQ, begin=a, n=b)
        nrec = s.validate()
        part, _ = D.fastq_filter(s, nrec, [])
        got = [x + y for x, y in zip(got, part.cpu().tolist()[:3])]
        got_gc += float(D.gc_from_counts(s.seq_len, s.gc, nrec).double().sum())
        del s, part
    assert got == want == [reads, 150 * reads, want[2]]
    assert abs(got_gc - want_gc) <= 1e-9 * reads
