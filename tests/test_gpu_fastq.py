"""GPU parity: the CUDA FASTQ path (through the C ABI) against the CPU oracle, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from tools import synth

import exb_testutil as util

pytestmark = pytest.mark.gpu


def _cols_equal(tab, orc):
    assert tab["__n_rows__"] == orc.n
    for name in orc.names:
        off, dat = orc.column(name)
        col = tab[name]
        assert np.array_equal(col.offsets.cpu().numpy(), off), name
        assert col.data.cpu().numpy().tobytes() == dat.tobytes(), name
    assert np.array_equal(tab["description"].valid.cpu().numpy().astype(bool), orc.desc_valid)


def _check_text(dev, text):
    """Full table + per-record statistics of `text` equal the oracle's (or both reject it)."""
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O

    buf = D.to_device(text, dev)
    try:
        want = O.parse_fastq(text)
    except O.OracleError as e:
        with pytest.raises(D.FormatError) as gi:
            D.fastq_table(buf)
        if e.msg in ("invalid name prefix", "invalid description prefix"):
            assert gi.value.pos == e.pos
        return None
    tab = D.fastq_table(buf)
    _cols_equal(tab, want)
    scan = D.fastq_scan_sync(buf)
    n = scan.validate()
    assert n == want.n
    seqs, quals = want.strings("sequence"), want.strings("quality_scores")
    assert scan.seq_len[:n].cpu().tolist() == [len(s) for s in seqs]
    assert scan.gc[:n].cpu().tolist() == [O.gc_count(s) for s in seqs]
    assert scan.qual_len[:n].cpu().tolist() == [len(q) for q in quals]
    assert scan.qsum[:n].cpu().tolist() == [O.quality_sum(q) for q in quals]
    return want


@pytest.mark.parametrize("name", ["test.fastq", "test2.fastq", "fastq/copy-a.fastq", "fastq/copy-b.fastq"])
def test_reference_fixtures(cuda_device, golden_dir, name):
    text = open(os.path.join(golden_dir, name), "rb").read()
    want = _check_text(cuda_device, text)
    assert want.n == 2  # test_fastq_scan.test:5-8


@pytest.mark.parametrize("seed,n,kw", [
    (1, 1, {}), (2, 7, {}), (3, 300, {}), (4, 300, dict(crlf=True)), (5, 300, dict(final_eol=False)),
    (6, 2000, dict(max_len=40)), (7, 50, dict(min_len=3000, max_len=40000)), (8, 500, dict(max_len=0)),
    (9, 64, dict(min_len=16384 - 40, max_len=16384 + 40)), (10, 3000, dict(crlf=True, final_eol=False, max_len=200)),
])
def test_random_records(cuda_device, seed, n, kw):
    text, _ = util.random_fastq(seed, n, **kw)
    want = _check_text(cuda_device, text)
    assert want is not None and want.n == n


def test_tile_edge_sweep(cuda_device):
    """Shift a record stream byte by byte across tile (16 KiB), warp-run (2 KiB) and thread-run (64 B) edges."""
    body, _ = util.random_fastq(11, 120, min_len=100, max_len=180)
    for pad in list(range(0, 70)) + [2047, 2048, 2049, 16383 - 60, 16384 - 59, 16384 - 1]:
        first = util.fastq_text([(b"pad", None, b"A" * pad, b"I" * pad)])
        _check_text(cuda_device, first + body)


@pytest.mark.parametrize("text", [
    b"", b"@a\nACGT\n+\nIIII", b"@a\nACGT\n+\nIIII\n", b"@a d\r\nACGT\r\n+\r\nIIII\r\n", b"@a\nACGT\n+\nIII\r",
    b"@\n\n+\n\n", b"@a \nA\n+a\n@\n@b\nC\n+\n+\n",
    b"@a\nACGT\n+\nIIII\n\n",           # trailing blank line: invalid name prefix
    b"a\nACGT\n+\nIIII\n",               # no '@'
    b"@a\nACGT\n-\nIIII\n",              # no '+'
    b"@a\nACGT\n+\n",                    # truncated
    b"@a\nACGT\n",                       # truncated
    b"@a\nACGT\n+\nIIII\n@b\nAC",        # truncated second record
    b"\n",
])
def test_edge_cases(cuda_device, text):
    _check_text(cuda_device, text)


def test_newline_dense_input(cuda_device):
    """Worst case for the per-line outputs: every record is 6 bytes."""
    _check_text(cuda_device, b"@\n\n+\n\n" * 20000)


def test_long_lines_span_many_tiles(cuda_device):
    import random
    rng = random.Random(5)
    recs = [(b"long%d" % i, b"d", util.rand_seq(rng, L), util.rand_qual(rng, L)) for i, L in enumerate([100000, 5, 70000, 16384, 32768])]
    _check_text(cuda_device, util.fastq_text(recs))


def _gen(dev, kind, n, **kw):
    from exon_duckdb_b200 import _lib, device as D
    p = synth.gen_params(kind, n, **kw)
    host = synth.gen_host(p)
    buf = synth.gen_device(p, dev)
    assert buf.cpu().numpy().tobytes() == host.tobytes(), "device and host generators must agree byte for byte"
    return buf, host.tobytes()


def test_generated_illumina_matches_oracle(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    buf, text = _gen(cuda_device, "illumina", 20000, seed=20)
    want = O.parse_fastq(text)
    _cols_equal(D.fastq_table(buf), want)


@pytest.mark.parametrize("op,thr", [(">", 30.0), (">=", 30.0), ("<", 31.5), (">", 21.3), ("<=", 30.02), ("=", 30.0), ("!=", 30.0)])
def test_mean_quality_filter_count(cuda_device, op, thr):
    """C2: SELECT COUNT(*) ... WHERE list_avg(quality_score_string_to_list(quality_scores)) <op> thr."""
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    buf, text = _gen(cuda_device, "illumina", 30000, seed=21)
    scan = D.fastq_scan_sync(buf)
    n = scan.validate()
    agg, pas = D.fastq_filter(scan, n, [("mean_quality", op, thr)], want_pass=True)
    want = O.parse_fastq(text)
    flags = [O.mean_quality_pass(q, op, thr) for q in want.strings("quality_scores")]
    assert pas.cpu().numpy().astype(bool).tolist() == flags
    got = agg.cpu().tolist()
    seqs = want.strings("sequence")
    assert got[0] == sum(flags)
    assert got[1] == sum(len(s) for s, f in zip(seqs, flags) if f)
    assert got[2] == sum(O.gc_count(s) for s, f in zip(seqs, flags) if f)
    assert (got[0], n, ) == O.fastq_count_mean_quality(text, op, thr)[:2]


def test_mean_quality_thresholds_that_need_the_x87_path(cuda_device):
    """Thresholds equal to an achievable mean: double rounding of the long double quotient decides."""
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    import random
    rng = random.Random(3)
    recs = []
    for i in range(4000):
        L = rng.choice([3, 7, 49, 150, 151, 1000])
        recs.append((b"r%d" % i, None, b"A" * L, util.rand_qual(rng, L, 33, 74)))
    text = util.fastq_text(recs)
    buf = D.to_device(text, cuda_device)
    scan = D.fastq_scan_sync(buf)
    n = scan.validate()
    quals = [r[3] for r in recs]
    for thr in [O.mean_quality(quals[5]), O.mean_quality(quals[77]), 20.0 + 1.0 / 3.0, float(np.nextafter(O.mean_quality(quals[9]), 100))]:
        for op in (">", ">=", "=", "<"):
            _, pas = D.fastq_filter(scan, n, [("mean_quality", op, thr)], want_pass=True)
            assert pas.cpu().numpy().astype(bool).tolist() == [O.mean_quality_pass(q, op, thr) for q in quals]


def test_combined_predicates_and_projection(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    text, _ = util.random_fastq(31, 5000, min_len=20, max_len=200, tricky=False)
    buf = D.to_device(text, cuda_device)
    preds = [("mean_quality", ">", 45.0), ("gc_content", ">=", 0.5), ("seq_len", ">", 50)]
    tab = D.fastq_table(buf, columns=["name", "sequence"], preds=preds)
    want = O.parse_fastq(text)
    keep = [i for i, (s, q) in enumerate(zip(want.strings("sequence"), want.strings("quality_scores")))
            if O.mean_quality_pass(q, ">", 45.0) and float(O.gc_content(s)) >= 0.5 and len(s) > 50]
    assert 0 < len(keep) < want.n
    assert tab["name"].to_pylist() == [want.strings("name")[i] for i in keep]
    assert tab["sequence"].to_pylist() == [want.strings("sequence")[i] for i in keep]


def test_gc_content_from_scan_counts_is_bit_exact(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    text, _ = util.random_fastq(41, 3000, max_len=400)
    buf = D.to_device(text, cuda_device)
    scan = D.fastq_scan_sync(buf)
    n = scan.validate()
    got = D.gc_from_counts(scan.seq_len, scan.gc, n).cpu().numpy()
    want = np.array([O.gc_content(s) for s in O.parse_fastq(text).strings("sequence")], dtype=np.float32)
    assert got.tobytes() == want.tobytes()  # <= 1 ulp is the bar; identical bits is what we get


def test_chained_ranges_equal_one_pass(cuda_device):
    """exb_fastq_scan chained over arbitrary tile-aligned ranges == one pass (the end-to-end engine relies on it)."""
    import torch
    from exon_duckdb_b200 import _lib, device as D
    text, _ = util.random_fastq(51, 4000, min_len=50, max_len=3000)
    buf = D.to_device(text, cuda_device)
    one = D.fastq_scan_sync(buf, _lib.F_LINES | _lib.F_SEQ | _lib.F_QUAL)
    n = one.validate()
    L = _lib.lib()
    cap = one.rec_cap
    mk = lambda dt, k: torch.zeros(k, dtype=dt, device=cuda_device)
    line, sl, gc, ql, qs = mk(torch.int32, 4 * cap), mk(torch.int32, cap), mk(torch.int32, cap), mk(torch.int32, cap), mk(torch.int32, cap)
    ws = [D.workspace(len(text) + 16, cuda_device) for _ in range(2)]
    cuts = [0, 16384, 16384 * 3, 16384 * 4, 16384 * 11, len(text)]
    cuts = sorted(set(c for c in cuts if c <= len(text)))
    prev = None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(len(cuts) - 1):
        w = ws[i & 1]
        D.check(L.exb_fastq_scan(buf.data_ptr(), cuts[i], cuts[i + 1], 1 if i == len(cuts) - 2 else 0, prev, D.UINT64_MAX, 7,
                                 line.data_ptr(), 4 * cap, 0, sl.data_ptr(), gc.data_ptr(), ql.data_ptr(), qs.data_ptr(), cap,
                                 w.data_ptr(), w.numel(), st))
        prev = C.c_void_p(w.data_ptr())
    res = D.fetch_result(ws[(len(cuts) - 2) & 1])
    assert res.total_lines == 4 * n and res.err_pos == _lib.NO_POS
    for a, b in ((line[:4 * n], one.line_end[:4 * n]), (sl[:n], one.seq_len[:n]), (gc[:n], one.gc[:n]), (ql[:n], one.qual_len[:n]),
                 (qs[:n], one.qsum[:n])):
        assert torch.equal(a, b)


@pytest.mark.parametrize("seed,n,kw", [(61, 3000, {}), (62, 500, dict(crlf=True, final_eol=False)), (63, 40, dict(min_len=3000, max_len=40000)),
                                         (64, 5000, dict(max_len=0)), (65, 1, {})])
def test_fused_scan_filter_equals_scan_then_filter(cuda_device, seed, n, kw):
    """exb_fastq_scan_filter (one kernel, nothing per record written) == exb_fastq_scan + exb_fastq_filter == oracle."""
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    text, _ = util.random_fastq(seed, n, **kw)
    buf = D.to_device(text, cuda_device)
    scan = D.fastq_scan_sync(buf)
    assert scan.validate() == n
    quals = O.parse_fastq(text).strings("quality_scores")
    for preds in ([("mean_quality", ">", 45.0)], [("mean_quality", "<=", 46.5), ("qual_len", ">", 10)], [], [("qual_len", "=", 0)]):
        want, _ = D.fastq_filter(scan, n, preds)
        got = D.fastq_scan_filter(buf, preds)
        assert got.validate() == n
        w, g = want.cpu().tolist(), got.agg.cpu().tolist()
        assert (g[0], g[3], g[4]) == (w[0], w[3], w[4]), preds
        ops = {"mean_quality": lambda q, op, v: O.mean_quality_pass(q, op, v),
               "qual_len": lambda q, op, v: {">": len(q) > v, "=": len(q) == v}[op]}
        assert g[0] == sum(all(ops[f](q, op, v) for f, op, v in preds) for q in quals)


def test_fused_scan_filter_generated_illumina(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    buf, text = _gen(cuda_device, "illumina", 30000, seed=20)
    c = D.fastq_scan_filter(buf, [("mean_quality", ">", 30.0)])
    assert c.validate() == 30000
    assert c.agg.cpu().tolist()[0] == O.fastq_count_mean_quality(text, ">", 30.0)[0]
    # reuse of the workspace / aggregate buffers across launches (what bench.py does every step)
    for _ in range(3):
        D.fastq_scan_filter(buf, [("mean_quality", ">", 30.0)], out=c)
    assert c.validate() == 30000
    assert c.agg.cpu().tolist()[0] == O.fastq_count_mean_quality(text, ">", 30.0)[0]


def test_fused_scan_filter_rejects_malformed(cuda_device):
    from exon_duckdb_b200 import device as D
    bad = b"@a\nACGT\n+\nIIII\n" * 300 + b"a\nACGT\n+\nIIII\n"
    with pytest.raises(D.FormatError) as e:
        D.fastq_scan_filter(D.to_device(bad, cuda_device), [("mean_quality", ">", 1.0)]).validate()
    assert e.value.pos == 15 * 300
    with pytest.raises(_lib_error()):
        D.fastq_scan_filter(D.to_device(bad, cuda_device), [("gc_content", ">", 0.5)])


def _lib_error():
    from exon_duckdb_b200 import _lib
    return _lib.ExonError
