"""GPU parity for the sequence scalar functions, the Arrow-stream reader (new_reader) and the host-buffer engine."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from tools import synth

import exb_testutil as util

pytestmark = pytest.mark.gpu


def _column(dev, strings):
    import torch
    from exon_duckdb_b200 import device as D
    off = np.zeros(len(strings) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in strings])
    data = D.to_device(b"".join(strings), dev)
    return D.Column(torch.from_numpy(off).to(dev), data)


def test_gc_content_column(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    rng = random.Random(1)
    seqs = [b"ATGC", b"ATGCGC", b"", b"GGGG", b"gcGC", b"GCN", b"ATGCGCA"]  # test_scalar_functions.test:5-28 + SURVEY 8c
    seqs += [util.rand_seq(rng, rng.choice([0, 1, 15, 16, 17, 31, 150, 1000, 70001]), b"ACGTNacgt") for _ in range(300)]
    got = D.gc_content(_column(cuda_device, seqs)).cpu().numpy()
    want = np.array([O.gc_content(s) for s in seqs], dtype=np.float32)
    assert got.tobytes() == want.tobytes()
    assert got[0] == np.float32(0.5) and got[3] == np.float32(1.0) and got[2] == np.float32(0.0)


def test_reverse_complement_and_complement(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    rng = random.Random(2)
    seqs = [b"ATCG", b"GGGG", b"ATGC", b"AACG", b"ATGCGC", b""] + [util.rand_seq(rng, rng.randint(0, 5000)) for _ in range(200)]
    col = _column(cuda_device, seqs)
    assert D.reverse_complement(col).to_pylist() == [O.reverse_complement(s) for s in seqs]
    assert D.complement(col).to_pylist() == [O.complement(s) for s in seqs]
    assert D.reverse_complement(col).to_pylist()[:2] == [b"CGAT", b"TTTT"]  # test_scalar_functions.test:41-46
    for bad in (b"acgt", b"ACGN", b"ATCGQ"):
        bad_col = _column(cuda_device, [b"ACGT" * 100, bad, b"GG"])
        for fn in (D.reverse_complement, D.complement):
            with pytest.raises(D.InvalidInput) as ei:
                fn(bad_col)
            with pytest.raises(O.InvalidInput) as eo:
                (O.reverse_complement if fn is D.reverse_complement else O.complement)(bad)
            assert str(ei.value) == str(eo.value)  # "Invalid character in sequence: <c>" names the first bad byte


def test_transcribe_reverse_transcribe_translate(cuda_device):
    """SURVEY 8f rank 3 (sequence_functions/module.cpp:168-360) on device columns vs the oracle, which is pinned on the
    reference's own functions (tests/golden/scalar_ref_vectors.json): values, and for the first offending row the
    reference's message -- rows in order, a row's length before its codons."""
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    rng = random.Random(5)
    dna = [b"ATCG", b"ATCGATCG", b""] + [util.rand_seq(rng, rng.randint(0, 3000)) for _ in range(300)]
    col = _column(cuda_device, dna)
    assert D.transcribe(col).to_pylist() == [O.transcribe(s) for s in dna]
    rna = [O.transcribe(s) for s in dna]
    assert D.reverse_transcribe(_column(cuda_device, rna)).to_pylist() == dna
    cod = [b"ATGCGC", b"", b"AAA"] + [util.rand_seq(rng, 3 * rng.randint(0, 1500)) for _ in range(300)]
    assert D.translate_dna_to_aa(_column(cuda_device, cod)).to_pylist() == [O.translate_dna_to_aa(s) for s in cod]
    for fn, ofn, bad in ((D.transcribe, O.transcribe, b"AUCG"), (D.reverse_transcribe, O.reverse_transcribe, b"ATCG"), (D.transcribe, O.transcribe, b"ACGn")):
        with pytest.raises(D.InvalidInput) as ei:
            fn(_column(cuda_device, [b"ACG" * 50, bad, b"GG"]))
        with pytest.raises(O.InvalidInput) as eo:
            ofn(bad)
        assert str(ei.value) == str(eo.value)
    # first offending row wins; inside a row the length is checked before the codons
    for rows_, first_bad in (([b"ATG", b"ATGNNN", b"ATGA"], 1), ([b"ATG", b"ATGA", b"NNN"], 1), ([b"ATGNNNA", b"AAA"], 0), ([b"atg"], 0),
                             ([b"AAA" * 700, b"AAANAAAAT"], 1)):
        with pytest.raises(D.InvalidInput) as ei:
            D.translate_dna_to_aa(_column(cuda_device, rows_))
        with pytest.raises(O.InvalidInput) as eo:
            O.translate_dna_to_aa(rows_[first_bad])
        assert str(ei.value) == str(eo.value), rows_


@pytest.mark.parametrize("mode", ["reverse_complement", "complement"])
def test_map_fused_into_the_gather(cuda_device, mode):
    """read_fastq -> reverse_complement(sequence) as one pass (exb_fastq_gather_map) equals gather + scalar function,
    rows of every length (pieces meeting inside 16-byte chunks, empty sequences), filtered or not, and names the first
    invalid byte exactly like the reference's scalar function."""
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    rng = random.Random(31)
    ref = O.reverse_complement if mode == "reverse_complement" else O.complement
    recs = []
    for i in range(3000):
        L = rng.choice([0, 1, 3, 15, 16, 17, 31, 150, rng.randint(0, 700)])
        seq = util.rand_seq(rng, L)
        recs.append(b"@r%d d\n" % i + seq + b"\n+\n" + bytes(rng.randint(35, 73) for _ in range(L)) + b"\n")
    text = b"".join(recs)
    buf = D.to_device(text, cuda_device)
    want = [ref(s) for s in O.parse_fastq(text).strings("sequence")]
    assert D.fastq_table(buf, columns=["sequence"], seq_map=mode)["sequence"].to_pylist() == want
    preds = [("mean_quality", ">", 20.0)]
    quals = O.parse_fastq(text).strings("quality_scores")
    keep = [w for w, q in zip(want, quals) if O.mean_quality_pass(q, ">", 20.0)]
    assert 0 < len(keep) < len(want)
    assert D.fastq_table(buf, columns=["sequence"], preds=preds, seq_map=mode)["sequence"].to_pylist() == keep
    bad = text + b"@bad x\nACGNT\n+\nIIIII\n@bad2\nAQ\n+\nII\n"  # two invalid bytes: the first one is reported
    with pytest.raises(D.InvalidInput) as ei:
        D.fastq_table(D.to_device(bad, cuda_device), columns=["sequence"], seq_map=mode)
    assert str(ei.value) == "Invalid character in sequence: N"


def test_quality_score_string_to_list(cuda_device):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O
    rng = random.Random(3)
    quals = [b"!'*5I~", b"IIII5555", bytes([0x80, 0xFF, 0x21])] + [util.rand_qual(rng, rng.randint(0, 3000), 0, 255) for _ in range(100)]
    col = _column(cuda_device, quals)
    got = D.quality_score_string_to_list(col).cpu().numpy()
    want = np.concatenate([O.quality_score_string_to_list(q) for q in quals])
    assert got.dtype == np.int32 and np.array_equal(got, want)
    assert got[:6].tolist() == [0, 6, 9, 20, 40, 93]


# ---------------------------------------------------------------- new_reader
class _Stream(C.Structure):
    _fields_ = [("get_schema", C.c_void_p), ("get_next", C.c_void_p), ("get_last_error", C.c_void_p),
                ("release", C.c_void_p), ("private_data", C.c_void_p)]


def _read(uri, fmt, filters=None, compression=None, batch=2048):
    import pyarrow as pa
    from exon_duckdb_b200 import _lib
    s = _Stream()
    r = _lib.lib().new_reader(C.byref(s), uri.encode(), batch, compression, fmt.encode(), filters)
    if r.error:
        raise RuntimeError(C.cast(r.error, C.c_char_p).value.decode())
    rd = pa.RecordBatchReader._import_from_c(C.addressof(s))
    batches = list(rd)
    assert all(0 < b.num_rows <= batch for b in batches)
    return pa.Table.from_batches(batches, schema=rd.schema)


def _rows(table):
    cols = [[None if v is None else v.encode("latin-1") for v in table.column(i).to_pylist()] for i in range(table.num_columns)]
    return list(zip(*cols))


def test_reader_reference_queries(cuda_device, golden_dir):
    from oracle import oracle as O
    fq = os.path.join(golden_dir, "test.fastq")
    t = _read(fq, "fastq")
    assert t.num_rows == 2  # test_fastq_scan.test:5-8
    assert _rows(t) == O.parse_fastq(open(fq, "rb").read()).rows()  # :34-41 incl. column order
    fa = os.path.join(golden_dir, "test.fasta")
    assert _read(fa, "fasta").num_rows == 2  # test_fasta_scan.test:5-8
    assert _read(fa, "fasta", b"id='a'").num_rows == 1  # :34-37 (FilterToString output for WHERE id = 'a')
    md = _read(os.path.join(golden_dir, "test.mixed-desc.fasta"), "fasta")
    assert md.column("description").to_pylist() == ["description", None]
    assert _read(os.path.join(golden_dir, "test.mixed-desc.fasta"), "fasta", b"description IS NULL").column("id").to_pylist() == ["b"]
    assert _read(os.path.join(golden_dir, "fastq") + "/", "fastq").num_rows == 4  # test_fastq_scan.test:64-68


def test_reader_gzip(cuda_device, tmp_path):
    import gzip
    from oracle import oracle as O
    text, _ = util.random_fastq(5, 500)
    p = tmp_path / "x.fastq.gz"
    with gzip.open(p, "wb") as f:
        f.write(text)
    assert _rows(_read(str(p), "fastq")) == O.parse_fastq(text).rows()          # auto-detect from ".gz"
    q = tmp_path / "y.fastq.gzip"
    q.write_bytes(p.read_bytes())
    assert _read(str(q), "fastq", compression=b"gzip").num_rows == 500          # explicit option


def test_reader_zstd(cuda_device, tmp_path, golden_dir, monkeypatch):
    """zstd input (test_fastq_scan.test:22-32): the reference's own .zst fixtures, a multi-block file
    written by another zstd implementation, concatenated frames, and a truncated file."""
    import pyarrow as pa
    from oracle import oracle as O
    assert _read(os.path.join(golden_dir, "test.fastq.zst"), "fastq").num_rows == 2
    assert _read(os.path.join(golden_dir, "test.fasta.zstd"), "fasta", compression=b"zstd").num_rows == 2
    text, _ = util.random_fastq(8, 4000, tricky=False)
    p = tmp_path / "x.fastq.zst"
    with pa.CompressedOutputStream(str(p), "zstd") as f:
        f.write(text)
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "100000")  # several blocks, records straddling their edges
    assert _rows(_read(str(p), "fastq")) == O.parse_fastq(text).rows()
    half = len(text) // 2
    cut = text.rfind(b"\n@", 0, half) + 1
    q = tmp_path / "two_frames.fastq.zst"
    with open(q, "wb") as out:  # two frames back to back decode as one stream
        for part in (text[:cut], text[cut:]):
            sink = pa.BufferOutputStream()
            with pa.CompressedOutputStream(sink, "zstd") as f:
                f.write(part)
            out.write(sink.getvalue().to_pybytes())
    assert _rows(_read(str(q), "fastq")) == O.parse_fastq(text).rows()
    t = tmp_path / "trunc.fastq.zst"
    t.write_bytes(p.read_bytes()[: p.stat().st_size // 2])
    with pytest.raises(Exception) as ei:
        _read(str(t), "fastq")
    assert "zstd" in str(ei.value)


@pytest.mark.parametrize("codec", ["bzip2", "xz"])
def test_reader_bzip2_and_xz(cuda_device, tmp_path, monkeypatch, codec):
    """FileCompressionType::BZIP2 / XZ of the reference's reader (datafusion 28), reachable through the `compression`
    option: one stream, two concatenated streams, several device blocks, and a truncated file."""
    import bz2
    import lzma
    from oracle import oracle as O
    comp = bz2.compress if codec == "bzip2" else lzma.compress
    text, _ = util.random_fastq(21, 4000, tricky=False)
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "150000")
    p = tmp_path / "x.fastq.cmp"
    p.write_bytes(comp(text))
    assert _rows(_read(str(p), "fastq", compression=codec.encode())) == O.parse_fastq(text).rows()
    cut = text.rfind(b"\n@", 0, len(text) // 2) + 1
    q = tmp_path / "two.fastq.cmp"
    q.write_bytes(comp(text[:cut]) + comp(text[cut:]))
    assert _rows(_read(str(q), "fastq", compression=codec.encode())) == O.parse_fastq(text).rows()
    fa, _ = util.random_fasta(22, 300, max_len=2000, tricky=False)
    f = tmp_path / "x.fasta.cmp"
    f.write_bytes(comp(fa))
    assert _rows(_read(str(f), "fasta", compression=(b"BZ2" if codec == "bzip2" else b"XZ"))) == O.parse_fasta(fa).rows()
    t = tmp_path / "trunc.fastq.cmp"
    t.write_bytes(p.read_bytes()[: p.stat().st_size // 2])
    with pytest.raises(Exception) as ei:
        _read(str(t), "fastq", compression=codec.encode())
    assert codec in str(ei.value) or "unexpected EOF" in str(ei.value)


@pytest.mark.parametrize("chunk", [None, 4096, 70000])
def test_reader_chunk_carry_fastq(cuda_device, tmp_path, monkeypatch, chunk):
    """Records straddling chunk edges are re-read with the next chunk; tiny chunks force that on every record."""
    from oracle import oracle as O
    if chunk:
        monkeypatch.setenv("EXON_B200_CHUNK_BYTES", str(chunk))
    text, _ = util.random_fastq(6, 3000, max_len=400, tricky=False)
    p = tmp_path / "a.fastq"
    p.write_bytes(text)
    want = O.parse_fastq(text)
    assert _rows(_read(str(p), "fastq")) == want.rows()
    got = _read(str(p), "fastq", b"mean_quality(quality_scores) > 45 AND name>='r2'")
    keep = [r for r in want.rows() if O.mean_quality_pass(r[3], ">", 45) and r[0] >= b"r2"]
    assert 0 < len(keep) < want.n and _rows(got) == keep


@pytest.mark.parametrize("chunk", [None, 8192, 100000])
def test_reader_chunk_carry_fasta(cuda_device, tmp_path, monkeypatch, chunk):
    from oracle import oracle as O
    if chunk:
        monkeypatch.setenv("EXON_B200_CHUNK_BYTES", str(chunk))
    text, _ = util.random_fasta(7, 800, max_len=3000, tricky=False)
    p = tmp_path / "a.fasta"
    p.write_bytes(text)
    want = O.parse_fasta(text)
    assert _rows(_read(str(p), "fasta")) == want.rows()
    got = _read(str(p), "fasta", b"gc_content(sequence) > 0.5 OR description IS NULL")
    keep = [r for r in want.rows() if float(O.gc_content(r[2])) > 0.5 or r[1] is None]
    assert 0 < len(keep) < want.n and _rows(got) == keep


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_reader_records_longer_than_a_block(cuda_device, tmp_path, monkeypatch, fmt):
    """The unconsumed tail of a chunk stays in HBM and the next block is appended behind it; a record that spans whole
    blocks makes the chunk grow (and the blocks double) until it holds a complete record.  Same rows either way."""
    from oracle import oracle as O
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "2048")
    p = tmp_path / ("a." + fmt)
    if fmt == "fastq":
        text, _ = util.random_fastq(11, 300, max_len=5000, min_len=10, tricky=False)
        want = O.parse_fastq(text)
    else:
        text, _ = util.random_fasta(12, 120, max_len=20000, tricky=False)
        want = O.parse_fasta(text)
    p.write_bytes(text)
    assert _rows(_read(str(p), fmt)) == want.rows()


def test_reader_batches_outlive_the_reader(cuda_device, tmp_path):
    """exb_batch views point into pinned result buffers owned by a pool that is shared with the batches: they stay
    valid after exb_reader_close (include/exon_b200.h, native reader contract)."""
    import ctypes as C
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    text, _ = util.random_fastq(13, 5000, tricky=False)
    p = tmp_path / "a.fastq"
    p.write_bytes(text)
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.exb_reader_open(str(p).encode(), b"fastq", None, 2048, None, 0xF, C.byref(h)))
    batches = []
    while True:
        b = _lib.Batch()
        _lib.check(L.exb_reader_next(h, C.byref(b)))
        if b.n_rows == 0:
            break
        batches.append(b)
    L.exb_reader_close(h)
    seqs = []
    for b in batches:
        off, data = b.cols[2].offsets, b.cols[2].data
        for i in range(b.n_rows):
            seqs.append(bytes(data[off[i]:off[i + 1]]))
        L.exb_batch_release(C.byref(b))
    assert seqs == O.parse_fastq(text).strings("sequence")


def test_reader_reports_malformed_input(cuda_device, tmp_path):
    p = tmp_path / "bad.fastq"
    p.write_bytes(b"@a\nACGT\n+\nIIII\nACGT\n")
    with pytest.raises(Exception) as e:
        _read(str(p), "fastq")
    assert "invalid FASTQ record at byte 15" in str(e.value)


# ---------------------------------------------------------------- host-buffer engine
@pytest.mark.parametrize("pinned", [False, True])
def test_engine_fastq_count_host(cuda_device, pinned):
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    L = _lib.lib()
    p = synth.gen_params("illumina", 40000, seed=20)
    n = synth.gen_size(p)
    if pinned:
        ptr = L.exb_host_alloc(n)
        host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,))
    else:
        host = np.empty(n, np.uint8)
    assert synth.lib().exb_gen_host(C.byref(p), host.ctypes.data, n) == 0
    preds, k = _lib.predicates([("mean_quality", ">", 30.0)])
    agg = (C.c_int64 * 8)()
    res = _lib.ScanResult()
    # 1 MiB chunks: 14 chained ranges, records straddle every edge
    _lib.check(L.exb_fastq_count_host(host.ctypes.data, n, preds, k, 1 << 20, 0, agg, C.byref(res)))
    want = O.fastq_count_mean_quality(host, ">", 30.0)
    assert (agg[0], agg[5]) == want[:2] and res.total_lines == 4 * 40000
    eng = C.c_void_p()
    _lib.check(L.exb_engine_create(0, 1 << 22, C.byref(eng)))
    for _ in range(2):
        agg2 = (C.c_int64 * 8)()
        _lib.check(L.exb_engine_fastq_count(eng, host.ctypes.data, n, preds, k, agg2, None))
        assert list(agg2)[:6] == list(agg)[:6]
    L.exb_engine_destroy(eng)
    if pinned:
        L.exb_host_free(ptr)
