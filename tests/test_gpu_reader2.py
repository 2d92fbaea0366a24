"""The native reader's round-2 surface (include/exon_b200.h: exb_reader_open2 / exb_reader_options): byte-range shards,
computed columns, device-built DuckDB string_t entries and validity bitmaps, several consumers on one reader, COUNT after
NEXT.  Everything is compared with the oracle on the same bytes (bit-exact: strings, lists, float32 / double values)."""
import ctypes as C
import struct
import threading

import numpy as np
import pytest

import exb_testutil as util

pytestmark = pytest.mark.gpu


def _open(path, fmt, filters=None, batch=2048, **kw):
    from exon_duckdb_b200 import _lib
    h = C.c_void_p()
    o = _lib.reader_options(**kw)
    _lib.check(_lib.lib().exb_reader_open2(str(path).encode(), fmt.encode(), None, batch, filters, C.byref(o), C.byref(h)))
    return h


def _string_t(ptr, n):
    """Decode n DuckDB string_t entries (string_type.hpp:19-60) the way DuckDB reads them."""
    raw = C.string_at(ptr, 16 * n)
    out = []
    for i in range(n):
        ln, = struct.unpack_from("<I", raw, 16 * i)
        if ln <= 12:
            s = raw[16 * i + 4:16 * i + 4 + ln]
            assert raw[16 * i + 4 + ln:16 * i + 16] == b"\0" * (12 - ln)  # zero padded: string_t compares the whole struct
        else:
            p, = struct.unpack_from("<Q", raw, 16 * i + 8)
            s = C.string_at(p, ln)
            assert raw[16 * i + 4:16 * i + 8] == s[:4]  # the 4-byte prefix
        out.append(s)
    return out


def _drain(h, ncols_total, want_strings=True):
    """All rows of a reader: per output column a python list (bytes / None / float / list of int)."""
    from exon_duckdb_b200 import _lib
    L = _lib.lib()
    cols = [[] for _ in range(ncols_total)]
    idx = []
    while True:
        b = _lib.Batch()
        _lib.check(L.exb_reader_next(h, C.byref(b)))
        if b.n_rows == 0:
            break
        n = b.n_rows
        idx.append(b.batch_index)
        for c in range(b.n_cols + b.n_computed):
            v = b.cols[c]
            if v.type == _lib.T_VARCHAR:
                if not (v.strings or v.offsets):
                    continue
                if v.strings:
                    vals = _string_t(v.strings, n)
                    if v.offsets:
                        assert vals == [bytes(v.data[v.offsets[i]:v.offsets[i + 1]]) for i in range(n)]
                else:
                    vals = [bytes(v.data[v.offsets[i]:v.offsets[i + 1]]) for i in range(n)]
            elif v.type == _lib.T_INT32_LIST:
                child = np.ctypeslib.as_array(C.cast(v.values, C.POINTER(C.c_int32)), shape=(max(v.n_values, 1),))[:v.n_values]
                vals = [child[v.list_entries[2 * i]:v.list_entries[2 * i] + v.list_entries[2 * i + 1]].tolist() for i in range(n)]
                assert sum(len(x) for x in vals) == v.n_values
            else:
                ct = {_lib.T_FLOAT: C.c_float, _lib.T_DOUBLE: C.c_double, _lib.T_INT64: C.c_int64}[v.type]
                vals = np.ctypeslib.as_array(C.cast(v.values, C.POINTER(ct)), shape=(n,)).copy().tolist()
                if v.type == _lib.T_FLOAT:
                    vals = [np.float32(x) for x in vals]
            if v.valid:
                if v.chunk_nulls == 0:
                    assert all(v.valid[i] for i in range(n))
                if v.valid_bits:
                    assert [(v.valid_bits[i >> 6] >> (i & 63)) & 1 for i in range(n)] == [1 if v.valid[i] else 0 for i in range(n)]
                vals = [x if v.valid[i] else None for i, x in enumerate(vals)]
            cols[c].extend(vals)
        L.exb_batch_release(C.byref(b))
    assert idx == list(range(len(idx)))
    return cols


def test_string_t_entries_and_validity_bitmaps(cuda_device, tmp_path):
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    text, _ = util.random_fastq(41, 7000, min_len=0, max_len=40, tricky=True)
    p = tmp_path / "s.fastq"
    p.write_bytes(text)
    ref = O.parse_fastq(text)
    for flags in (_lib.RD_STRING_T, _lib.RD_STRING_T | _lib.RD_NO_OFFSETS):
        h = _open(p, "fastq", column_mask=0xF, flags=flags)
        cols = _drain(h, 4)
        _lib.lib().exb_reader_close(h)
        assert cols[0] == ref.strings("name") and cols[2] == ref.strings("sequence") and cols[3] == ref.strings("quality_scores")
        assert cols[1] == [d if v else None for d, v in zip(ref.strings("description"), ref.desc_valid)]


def test_computed_columns_match_the_oracle(cuda_device, tmp_path, monkeypatch):
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "180000")
    text, _ = util.random_fastq(43, 5000, min_len=0, max_len=200, tricky=False)
    p = tmp_path / "c.fastq"
    p.write_bytes(text)
    ref = O.parse_fastq(text)
    seqs, quals = ref.strings("sequence"), ref.strings("quality_scores")
    comp = [(_lib.C_GC_CONTENT, 0), (_lib.C_SEQ_MAP, _lib.MAP_REVERSE_COMPLEMENT), (_lib.C_QUALITY_LIST, 0), (_lib.C_MEAN_QUALITY, 0),
            (_lib.C_SEQ_MAP, _lib.MAP_COMPLEMENT), (_lib.C_SEQ_LENGTH, 0), (_lib.C_QUAL_LENGTH, 0)]
    for filt, keep in ((None, list(range(ref.n))),
                       (b"mean_quality(quality_scores)>79.5 AND name>='r2'", [i for i in range(ref.n) if O.mean_quality_pass(quals[i], ">", 79.5) and ref.strings("name")[i] >= b"r2"])):
        h = _open(p, "fastq", filters=filt, column_mask=0x1, flags=_lib.RD_STRING_T, computed=comp)  # only `name` of the file columns
        cols = _drain(h, 4 + len(comp))
        _lib.lib().exb_reader_close(h)
        assert cols[0] == [ref.strings("name")[i] for i in keep] and cols[2] == [] and cols[3] == []
        assert cols[4] == [O.gc_content(seqs[i]) for i in keep]
        assert cols[5] == [O.reverse_complement(seqs[i]) for i in keep]
        assert cols[6] == [[c - 33 for c in quals[i]] for i in keep]
        want_mean = [float(np.longdouble(sum(c - 33 for c in quals[i])) / np.longdouble(len(quals[i]))) if quals[i] else None for i in keep]
        assert cols[7] == want_mean
        assert cols[8] == [O.complement(seqs[i]) for i in keep]
        assert cols[9] == [len(seqs[i]) for i in keep] and cols[10] == [len(quals[i]) for i in keep]


def test_computed_columns_fasta(cuda_device, tmp_path, monkeypatch):
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "100000")
    text, _ = util.random_fasta(47, 150, min_len=0, max_len=6000, tricky=False)
    p = tmp_path / "c.fasta"
    p.write_bytes(text)
    ref = O.parse_fasta(text)
    seqs = ref.strings("sequence")
    comp = [(_lib.C_GC_CONTENT, 0), (_lib.C_SEQ_MAP, _lib.MAP_TRANSCRIBE), (_lib.C_SEQ_LENGTH, 0)]
    h = _open(p, "fasta", column_mask=0x1, flags=_lib.RD_STRING_T | _lib.RD_NO_OFFSETS, computed=comp)
    cols = _drain(h, 3 + len(comp))
    _lib.lib().exb_reader_close(h)
    assert cols[0] == ref.strings("id")
    assert cols[3] == [O.gc_content(s) for s in seqs]
    assert cols[4] == [s.replace(b"T", b"U") for s in seqs]
    assert cols[5] == [len(s) for s in seqs]
    with pytest.raises(_lib.ExonError):  # quality columns do not exist in FASTA
        _open(p, "fasta", computed=[(_lib.C_MEAN_QUALITY, 0)])


def test_map_column_reports_the_offending_byte(cuda_device, tmp_path):
    from exon_duckdb_b200 import _lib
    text = util.fastq_text([(b"a", None, b"ACGT", b"IIII"), (b"b", None, b"ACGTNACGT", b"IIIIIIIII")])
    p = tmp_path / "n.fastq"
    p.write_bytes(text)
    h = _open(p, "fastq", column_mask=0, computed=[(_lib.C_SEQ_MAP, _lib.MAP_REVERSE_COMPLEMENT)])
    b = _lib.Batch()
    rc = _lib.lib().exb_reader_next(h, C.byref(b))
    assert rc == _lib.ERR_INVALID_CHAR and b"Invalid character in sequence: N" in _lib.lib().exb_last_error()
    _lib.lib().exb_reader_close(h)


def _rows_of(path, fmt, ncols, **kw):
    from exon_duckdb_b200 import _lib
    h = _open(path, fmt, column_mask=(1 << ncols) - 1, **kw)
    cols = _drain(h, ncols)
    _lib.lib().exb_reader_close(h)
    return list(zip(*cols)) if cols[0] else []


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_byte_range_shards_swept_over_every_offset(cuda_device, tmp_path, fmt):
    """N-shard output equals 1-shard output byte for byte, with the shard boundary swept across every byte offset of
    a small file (the 'fake cluster').  A record belongs to the shard that holds its first byte."""
    from oracle import oracle as O
    if fmt == "fastq":
        text, _ = util.random_fastq(51, 7, min_len=1, max_len=30, tricky=False)
        want = O.parse_fastq(text).rows()
        ncols = 4
    else:
        text, _ = util.random_fasta(53, 6, min_len=0, max_len=90, tricky=False)
        want = O.parse_fasta(text).rows()
        ncols = 3
    p = tmp_path / ("sweep." + fmt)
    p.write_bytes(text)
    n = len(text)
    assert _rows_of(p, fmt, ncols) == want
    for cut in range(1, n):
        a = _rows_of(p, fmt, ncols, range_lo=0, range_hi=cut)
        b = _rows_of(p, fmt, ncols, range_lo=cut, range_hi=0)
        assert a + b == want, cut


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_byte_range_shards_of_a_larger_file(cuda_device, tmp_path, fmt, monkeypatch):
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "150000")
    if fmt == "fastq":
        text, _ = util.random_fastq(57, 9000, min_len=0, max_len=150, tricky=False)
        ref, ncols = O.parse_fastq(text), 4
    else:
        text, _ = util.random_fasta(59, 400, min_len=0, max_len=5000, tricky=False)
        ref, ncols = O.parse_fasta(text), 3
    want = ref.rows()
    p = tmp_path / ("big." + fmt)
    p.write_bytes(text)
    n = len(text)
    for shards in (2, 5, 8):
        got, counted = [], 0
        for k in range(shards):
            lo, hi = n * k // shards, (0 if k == shards - 1 else n * (k + 1) // shards)
            got += _rows_of(p, fmt, ncols, range_lo=lo, range_hi=hi, flags=_lib.RD_STRING_T)
            h = _open(p, fmt, column_mask=0, range_lo=lo, range_hi=hi)
            c = C.c_int64()
            _lib.check(_lib.lib().exb_reader_count(h, C.byref(c)))
            _lib.lib().exb_reader_close(h)
            counted += c.value
        assert got == want and counted == len(want)


def test_a_wrong_fastq_cut_is_an_error_not_a_wrong_answer(cuda_device, tmp_path):
    """'@' and '+' are legal quality characters: a file built so that the resync rule picks a quality line as the record start
    makes the shard that ENDS there fail its own line-count check (include/exon_b200.h, exb_reader_options)."""
    from exon_duckdb_b200 import _lib
    # every quality line starts with '@' and every sequence line with '+': hypotheses "header" and "quality" both fit
    recs = [(b"r%d" % i, None, b"+ACGT", b"@IIII") for i in range(40)]
    text = util.fastq_text(recs)
    p = tmp_path / "adv.fastq"
    p.write_bytes(text)
    first_q = text.index(b"@IIII")  # cut right at a quality line: it looks like a header followed, two lines down, by '+ACGT'
    outcomes = []
    for lo, hi in ((0, first_q), (first_q, 0)):
        try:
            outcomes.append(_rows_of(p, "fastq", 4, range_lo=lo, range_hi=hi))
        except _lib.ExonError as e:
            outcomes.append(str(e))
    assert any(isinstance(o, str) and "byte-range shard" in o for o in outcomes), outcomes


def test_several_consumers_share_one_reader(cuda_device, tmp_path, monkeypatch):
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "200000")
    text, _ = util.random_fastq(61, 30000, min_len=10, max_len=60, tricky=False)
    p = tmp_path / "m.fastq"
    p.write_bytes(text)
    want = O.parse_fastq(text).strings("name")
    L = _lib.lib()
    h = _open(p, "fastq", column_mask=0x1, flags=_lib.RD_STRING_T)
    got = {}
    lock = threading.Lock()

    def work():
        while True:
            b = _lib.Batch()
            _lib.check(L.exb_reader_next(h, C.byref(b)))
            if b.n_rows == 0:
                return
            names = _string_t(b.cols[0].strings, b.n_rows)
            with lock:
                assert b.batch_index not in got
                got[b.batch_index] = names
            L.exb_batch_release(C.byref(b))

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    L.exb_reader_close(h)
    assert sorted(got) == list(range(len(got)))
    assert [x for k in sorted(got) for x in got[k]] == want


def test_count_after_next_counts_the_rest(cuda_device, tmp_path, monkeypatch):
    """exb_reader_count "consumes the rest of the stream" also when batches were already handed out (round-1 advisor finding:
    it used to spin forever on the unread rows of the current chunk)."""
    from exon_duckdb_b200 import _lib
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", "100000")
    text, _ = util.random_fastq(67, 12000, min_len=10, max_len=60, tricky=False)
    p = tmp_path / "k.fastq"
    p.write_bytes(text)
    L = _lib.lib()
    h = _open(p, "fastq", column_mask=0xF)
    seen = 0
    for _ in range(2):
        b = _lib.Batch()
        _lib.check(L.exb_reader_next(h, C.byref(b)))
        seen += b.n_rows
        L.exb_batch_release(C.byref(b))
    c = C.c_int64()
    _lib.check(L.exb_reader_count(h, C.byref(c)))
    done, total = C.c_int64(), C.c_int64()
    _lib.check(L.exb_reader_progress(h, C.byref(done), C.byref(total)))
    L.exb_reader_close(h)
    assert seen + c.value == 12000
    assert total.value == len(text) and done.value == len(text)


def test_gzip_input_that_is_not_gzip_or_is_truncated_fails(cuda_device, tmp_path):
    """The reference's GzipDecoder rejects both; gzread alone would pass plain text through and stop silently at a truncation."""
    import gzip
    from exon_duckdb_b200 import _lib
    text, _ = util.random_fastq(71, 3000, min_len=50, max_len=150, tricky=False)
    plain = tmp_path / "plain.fastq.gz"
    plain.write_bytes(text)
    z = gzip.compress(text)
    cut = tmp_path / "cut.fastq.gz"
    cut.write_bytes(z[:len(z) // 2])
    for path in (plain, cut):
        h = C.c_void_p()
        _lib.check(_lib.lib().exb_reader_open(str(path).encode(), b"fastq", None, 2048, None, 0xF, C.byref(h)))
        c = C.c_int64()
        rc = _lib.lib().exb_reader_count(h, C.byref(c))
        msg = _lib.lib().exb_last_error()
        _lib.lib().exb_reader_close(h)
        assert rc != 0 and b"gzip" in msg, (path, rc, msg)


def test_second_scan_reads_the_registered_page_cache(cuda_device, monkeypatch):
    """A memory-backed file that was scanned once is registered with CUDA in the background (its page cache becomes pinned
    memory); the next scans DMA from it instead of copying through pinned blocks.  Rows, borrowed string_t entries, the
    fused COUNT and byte-range shards must be identical on both paths."""
    import os
    import time
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    from tools import synth
    if not os.path.isdir("/dev/shm") or not os.access("/dev/shm", os.W_OK):
        pytest.skip("no tmpfs to put a memory-backed file on")
    L = _lib.lib()
    monkeypatch.setenv("EXON_B200_CHUNK_BYTES", str(3 << 20))  # several blocks, records straddling every edge
    monkeypatch.setenv("EXON_B200_REGISTER_PIECE_MB", "4")     # the mapping is pinned in pieces; 3 MiB blocks cross their edges
    text = synth.gen_host(synth.gen_params("illumina", 60000, seed=77)).tobytes()
    assert len(text) > (16 << 20)
    path = "/dev/shm/exb_test_registered_%d.fastq" % os.getpid()
    with open(path, "wb") as f:
        f.write(text)
    try:
        ref = O.parse_fastq(text)
        want_pass = O.fastq_count_mean_quality(text, ">", 30.0)[0]
        filt = b"mean_quality(quality_scores)>30.0"

        def count(**kw):
            h = _open(path, "fastq", filters=filt, column_mask=0, **kw)
            c = C.c_int64()
            _lib.check(L.exb_reader_count(h, C.byref(c)))
            direct = L.exb_reader_io_path(h)
            L.exb_reader_close(h)
            return c.value, direct

        assert L.exb_file_cache_state(path.encode()) == 0
        assert count() == (want_pass, 0)                      # first scan: the copy path; triggers the registration
        t0 = time.time()
        while L.exb_file_cache_state(path.encode()) == 1 and time.time() - t0 < 20:
            time.sleep(0.01)
        if L.exb_file_cache_state(path.encode()) != 2:
            pytest.skip("this kernel / file system does not let CUDA pin page-cache pages")
        assert count() == (want_pass, 1)                      # fused COUNT by DMA from the page cache
        assert count(flags=_lib.RD_COPY_IO) == (want_pass, 0)  # the flag keeps a reader on the copy path
        for flags in (_lib.RD_STRING_T | _lib.RD_NO_OFFSETS, 0):  # borrowed strings point into the mapping / gathered columns
            h = _open(path, "fastq", column_mask=0xF, flags=flags)
            cols = _drain(h, 4)
            assert L.exb_reader_io_path(h) == 1
            L.exb_reader_close(h)
            assert cols[0] == ref.strings("name") and cols[2] == ref.strings("sequence") and cols[3] == ref.strings("quality_scores")
        # byte-range shards (cuts inside records) of the registered file
        n = len(text)
        got = []
        for lo, hi in ((0, n // 3 + 5), (n // 3 + 5, 2 * n // 3 + 77), (2 * n // 3 + 77, n)):
            h = _open(path, "fastq", column_mask=0x1, flags=_lib.RD_STRING_T, range_lo=lo, range_hi=hi)
            got += _drain(h, 4)[0]
            L.exb_reader_close(h)
        assert got == ref.strings("name")
    finally:
        os.unlink(path)
