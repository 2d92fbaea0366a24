"""SURVEY 8(f) rank 1, BGZF input inflated on the device (exon_duckdb_b200/csrc/inflate.cu).

CPU tier: the DEFLATE / CRC-32 core the kernel runs (inflate_core.cuh) is compiled for the host with one lane
(oracle/libinflate_host.so, test infrastructure) and compared with zlib -- the oracle of this row, the algorithm lives in a
third-party dependency (zlib, RFC 1951 / 1952) -- over stored, fixed and dynamic blocks, every compression level and
strategy, long matches, overlapping matches, code words longer than the index tables, and corrupted streams.
GPU tier: the kernel itself, and the reader on bgzip'ed FASTQ / FASTA, against the same oracle."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

import exb_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host(small=False):
    so = os.path.join(ROOT, "oracle", "libinflate_host_small.so" if small else "libinflate_host.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libinflate_host.so"])
    lib = C.CDLL(so)
    lib.ifl_host_inflate.restype = C.c_int
    lib.ifl_host_inflate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
    lib.ifl_host_crc32.restype = C.c_uint32
    lib.ifl_host_crc32.argtypes = [C.c_char_p, C.c_int]
    return lib


raw_deflate, bgzf_bytes = util.raw_deflate, util.bgzf_bytes


def _payloads():
    rng = np.random.default_rng(5)
    fq, _ = util.random_fastq(3, 400, min_len=50, max_len=150, tricky=False)
    fa, _ = util.random_fasta(4, 30, min_len=100, max_len=3000, tricky=False)
    yield "empty", b""
    yield "one byte", b"x"
    yield "fastq", fq[:65000]
    yield "fasta", fa[:65000]
    yield "zeros", bytes(65536)
    yield "run of 3", b"abc" * 20000
    yield "random bytes", rng.integers(0, 256, 60000, dtype=np.uint8).tobytes()
    yield "random bases", rng.choice(np.frombuffer(b"ACGT", np.uint8), 65536).tobytes()
    # a skewed alphabet: code words longer than the 10-bit literal table and the 8-bit distance table
    p = np.array([2.0 ** -(i // 6) for i in range(256)])
    yield "skewed", rng.choice(256, 65000, p=p / p.sum()).astype(np.uint8).tobytes()
    words = [rng.integers(0, 256, int(rng.integers(3, 300)), dtype=np.uint8).tobytes() for _ in range(300)]
    yield "long matches far apart", b"".join(words[int(i)] for i in rng.integers(0, 300, 600))[:65536]


@pytest.mark.parametrize("small_tables", [False, True])
def test_core_matches_zlib_on_every_block_type(small_tables):
    """small_tables: a build whose index tables hold 4 / 2 bits, so that nearly every code word is decoded bit by bit from the
    canonical arrays -- the path code words longer than 10 / 8 bits take in the product build."""
    lib = _host(small_tables)
    n_cases = 0
    for name, data in _payloads():
        for level in (0, 1, 3, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                z = raw_deflate(data, level, strategy)
                for mis in range(4) if n_cases % 7 == 0 else (n_cases & 3,):
                    out = C.create_string_buffer(len(data) + 8)
                    rc = lib.ifl_host_inflate(z, len(z), out, len(data), mis)
                    assert rc == 0, (name, level, strategy, mis, rc)
                    assert out.raw[:len(data)] == data, (name, level, strategy)
                n_cases += 1
        assert lib.ifl_host_crc32(data, len(data)) == zlib.crc32(data), name
    # small memLevel: many short dynamic blocks in one stream
    fq, _ = util.random_fastq(9, 300, min_len=100, max_len=150, tricky=False)
    z = raw_deflate(fq[:65000], 6, mem=1)
    out = C.create_string_buffer(65000)
    assert lib.ifl_host_inflate(z, len(z), out, len(fq[:65000]), 1) == 0 and out.raw[:len(fq[:65000])] == fq[:65000]
    for n in (0, 1, 31, 32, 33, 1000, 65535, 65536):
        d = bytes((i * 7 + 3) & 255 for i in range(n))
        assert lib.ifl_host_crc32(d, n) == zlib.crc32(d), n


def test_core_rejects_corrupt_streams():
    lib = _host()
    fq, _ = util.random_fastq(11, 300, min_len=100, max_len=150, tricky=False)
    data = fq[:50000]
    z = raw_deflate(data, 6)
    out = C.create_string_buffer(len(data) + 8)
    assert lib.ifl_host_inflate(z, len(z), out, len(data) - 1, 0) != 0      # trailer says fewer bytes
    assert lib.ifl_host_inflate(z, len(z), out, len(data) + 1, 0) != 0      # ... more bytes
    assert lib.ifl_host_inflate(z, len(z) // 2, out, len(data), 0) != 0     # payload cut short
    rng = np.random.default_rng(1)
    detected = 0
    for _ in range(200):   # flipped bits: an error, or output that differs (which the CRC-32 then catches); never a crash
        b = bytearray(z)
        i = int(rng.integers(0, len(b)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        rc = lib.ifl_host_inflate(bytes(b), len(b), out, len(data), 0)
        if rc != 0 or lib.ifl_host_crc32(out.raw[:len(data)], len(data)) != zlib.crc32(data):
            detected += 1
    assert detected == 200
    assert lib.ifl_host_inflate(b"\x07\x00", 2, out, 0, 0) != 0               # reserved block type


def test_bgzf_writer_of_the_tests_is_valid_gzip():
    """The BGZF images the tests build are what zlib / gzip read back (multi-member gzip)."""
    import gzip
    fq, _ = util.random_fastq(13, 2000, min_len=50, max_len=150, tricky=False)
    img = bgzf_bytes(fq, sizes=[65280, 1, 777, 30000])
    assert gzip.decompress(img) == fq


# ------------------------------------------------------------------------------------------------ GPU tier
def _corrupt(img, member, what):
    """Damage one member of a BGZF image in place (sizes unchanged)."""
    from exon_duckdb_b200 import device as D
    tab, n, _, _ = D.bgzf_index(img)
    # member k's payload starts at in_off (relative to offset 0 here)
    b = bytearray(img)
    blk = tab[member]
    if what == "payload":
        b[blk.in_off + blk.clen // 2] ^= 0x10
    elif what == "crc":
        b[blk.in_off + blk.clen] ^= 0xFF
    return bytes(b)


@pytest.mark.gpu
def test_kernel_inflates_every_block_type(cuda_device):
    from exon_duckdb_b200 import device as D
    pieces = []
    images = []
    for name, data in _payloads():
        for level, strategy in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            pieces.append(data[:65000])
            images.append(bgzf_bytes(data[:65000], block=65000, level=level, strategy=strategy, eof_marker=False))
    img = b"".join(images) + bgzf_bytes(b"", eof_marker=False)
    want = b"".join(pieces)
    got = D.bgzf_inflate(img, cuda_device).cpu().numpy().tobytes()
    assert len(got) == len(want)
    assert got == want
    # members of every size class, several to a warp's lifetime
    fq, _ = util.random_fastq(21, 6000, min_len=50, max_len=150, tricky=False)
    img = bgzf_bytes(fq, sizes=[65280, 1, 2, 777, 30000, 31, 33, 4096])
    assert D.bgzf_inflate(img, cuda_device).cpu().numpy().tobytes() == fq


@pytest.mark.gpu
def test_kernel_reports_the_first_corrupt_member(cuda_device):
    from exon_duckdb_b200 import _lib, device as D
    fq, _ = util.random_fastq(23, 4000, min_len=100, max_len=150, tricky=False)
    img = bgzf_bytes(fq, block=20000)
    for member, what in ((3, "crc"), (5, "payload"), (0, "payload")):
        with pytest.raises(_lib.ExonError) as e:
            D.bgzf_inflate(_corrupt(img, member, what), cuda_device)
        assert "corrupt gzip stream" in str(e.value) and ("block %d" % member) in str(e.value), str(e.value)
    two = _corrupt(_corrupt(img, 7, "crc"), 2, "crc")
    with pytest.raises(_lib.ExonError) as e:
        D.bgzf_inflate(two, cuda_device)
    assert "block 2" in str(e.value)
    with pytest.raises(_lib.ExonError):     # not BGZF at all
        D.bgzf_index(b"\x1f\x8b\x08\x00" + bytes(30))


def _open(path, fmt, filters=None, **kw):
    from exon_duckdb_b200 import _lib
    h = C.c_void_p()
    o = _lib.reader_options(**kw)
    _lib.check(_lib.lib().exb_reader_open2(str(path).encode(), fmt.encode(), None, 2048, filters, C.byref(o), C.byref(h)))
    return h


def _rows(path, fmt, ncols):
    """All rows through the reader's Arrow-style columns."""
    from exon_duckdb_b200 import _lib
    L = _lib.lib()
    h = _open(path, fmt, column_mask=(1 << ncols) - 1)
    cols = [[] for _ in range(ncols)]
    while True:
        b = _lib.Batch()
        _lib.check(L.exb_reader_next(h, C.byref(b)))
        if b.n_rows == 0:
            break
        for c in range(ncols):
            v = b.cols[c]
            vals = [bytes(v.data[v.offsets[i]:v.offsets[i + 1]]) for i in range(b.n_rows)]
            if v.valid:
                vals = [x if v.valid[i] else None for i, x in enumerate(vals)]
            cols[c].extend(vals)
        L.exb_batch_release(C.byref(b))
    L.exb_reader_close(h)
    return cols


def _count(path, fmt, filters=None):
    from exon_duckdb_b200 import _lib
    L = _lib.lib()
    h = _open(path, fmt, filters=filters, column_mask=0)
    c = C.c_int64()
    rc = L.exb_reader_count(h, C.byref(c))
    msg = L.exb_last_error()
    L.exb_reader_close(h)
    return rc, c.value, msg


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [0, 150000, 4096])
def test_reader_on_bgzf_fastq_and_fasta(cuda_device, tmp_path, monkeypatch, chunk):
    """bgzip'ed input through the reader: rows, COUNT(*) (fused chained scan: members end anywhere, ranges on multiples of
    16) and the mean-quality filter, identical to the oracle on the text and to the streaming zlib path."""
    from oracle import oracle as O
    if chunk:
        # many blocks: tails, carried records; at 4096 (16 KiB of text per block) every member is a block of its own, and the
        # 5-byte and 1-byte members are blocks below the 16 bytes a chained range of the fused COUNT scan needs
        monkeypatch.setenv("EXON_B200_CHUNK_BYTES", str(chunk))
    fq, _ = util.random_fastq(31, 9000, min_len=20, max_len=150, tricky=False)
    fa, _ = util.random_fasta(33, 300, min_len=10, max_len=4000, tricky=False)
    pq = tmp_path / "r.fastq.gz"
    pa = tmp_path / "r.fasta.gz"
    pq.write_bytes(bgzf_bytes(fq, sizes=[65280, 5, 30011, 64000, 1]))
    pa.write_bytes(bgzf_bytes(fa, block=50000))
    rq, ra = O.parse_fastq(fq), O.parse_fasta(fa)
    want_pass = O.fastq_count_mean_quality(fq, ">", 30.0)[0]
    for bgzf in ("1", "0"):
        monkeypatch.setenv("EXON_B200_BGZF", bgzf)
        cq = _rows(pq, "fastq", 4)
        assert cq[0] == rq.strings("name") and cq[2] == rq.strings("sequence") and cq[3] == rq.strings("quality_scores")
        ca = _rows(pa, "fasta", 3)
        assert ca[0] == ra.strings("id") and ca[2] == ra.strings("sequence")
        assert _count(pq, "fastq")[:2] == (0, rq.n)
        assert _count(pq, "fastq", b"mean_quality(quality_scores)>30.0")[:2] == (0, want_pass)
        assert _count(pa, "fasta")[:2] == (0, ra.n)
    # empty text: the EOF marker alone
    pe = tmp_path / "e.fastq.gz"
    pe.write_bytes(bgzf_bytes(b""))
    monkeypatch.setenv("EXON_B200_BGZF", "1")
    assert _count(pe, "fastq")[:2] == (0, 0)


@pytest.mark.gpu
def test_reader_reports_a_corrupt_bgzf_member(cuda_device, tmp_path):
    fq, _ = util.random_fastq(37, 5000, min_len=100, max_len=150, tricky=False)
    img = bgzf_bytes(fq, block=40000)
    for what in ("crc", "payload"):
        p = tmp_path / ("bad_%s.fastq.gz" % what)
        p.write_bytes(_corrupt(img, 4, what))
        for filt in (None, b"mean_quality(quality_scores)>30.0"):
            rc, _, msg = _count(p, "fastq", filt)
            assert rc != 0 and b"corrupt gzip stream" in msg and b"block 4" in msg, (what, filt, rc, msg)
    p = tmp_path / "cut.fastq.gz"
    p.write_bytes(img[:len(img) // 2])
    rc, _, msg = _count(p, "fastq")
    assert rc != 0 and b"BGZF" in msg, msg


@pytest.mark.gpu
def test_rows_of_a_corrupt_bgzf_file_are_an_error(cuda_device, tmp_path):
    from exon_duckdb_b200 import _lib
    fa, _ = util.random_fasta(41, 200, min_len=100, max_len=3000, tricky=False)
    p = tmp_path / "bad.fasta.gz"
    p.write_bytes(_corrupt(bgzf_bytes(fa, block=30000), 2, "payload"))
    with pytest.raises(_lib.ExonError) as e:
        _rows(p, "fasta", 3)
    assert "corrupt gzip stream" in str(e.value) and "block 2" in str(e.value)


def test_bgzf_compress_host_round_trips():
    """The writers' gzip sink (exb_bgzf_compress_host): BGZF members any gzip reader reads, whose member table the index walk
    finds, and whose payloads the inflate core decodes.  Host code of the product library: no GPU needed."""
    import gzip
    from exon_duckdb_b200 import _lib
    L = _lib.lib()
    host = _host()
    rng = np.random.default_rng(9)
    fq, _ = util.random_fastq(43, 3000, min_len=50, max_len=150, tricky=False)
    for data in (b"", b"x", fq, rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(), bytes(65280 * 2)):
        for threads in (1, 5):
            cap = L.exb_bgzf_compress_bound(len(data))
            out = C.create_string_buffer(cap)
            n = C.c_int64()
            _lib.check(L.exb_bgzf_compress_host(data, len(data), 6, threads, 1, out, cap, C.byref(n)))
            img = out.raw[:n.value]
            assert gzip.decompress(img) == data
            assert L.exb_bgzf_probe_host(img, len(img)) == 1
            tab = (_lib.BgzfBlock * (len(data) // 65280 + 3))()
            nb, nxt, outb = C.c_int64(), C.c_int64(), C.c_int64()
            _lib.check(L.exb_bgzf_index_host(img, len(img), 0, 1 << 40, tab, len(tab), C.byref(nb), C.byref(nxt), C.byref(outb)))
            assert nxt.value == len(img) and outb.value == len(data) and nb.value == (len(data) + 65279) // 65280 + 1
            got = b""
            for k in range(nb.value):
                b = tab[k]
                piece = C.create_string_buffer(b.isize + 8)
                assert host.ifl_host_inflate(img[b.in_off:b.in_off + b.clen], b.clen, piece, b.isize, k & 3) == 0
                assert zlib.crc32(piece.raw[:b.isize]) == b.crc32
                got += piece.raw[:b.isize]
            assert got == data


def test_bgzf_member_walk_edge_cases():
    """exb_bgzf_probe_host / exb_bgzf_index_host (host code): other extra subfields next to BC, members that are not BGZF
    (plain gzip, FNAME set), truncated files, size limits, max_out_bytes / max_blocks cuts."""
    from exon_duckdb_b200 import _lib
    L = _lib.lib()

    def member(piece, extra_before=b"", extra_after=b"", flags=4):
        z = raw_deflate(piece)
        xlen = len(extra_before) + 6 + len(extra_after)
        bsize = 12 + xlen + len(z) + 8
        hdr = b"\x1f\x8b\x08" + bytes([flags]) + b"\0\0\0\0\x00\xff" + xlen.to_bytes(2, "little") + extra_before + b"BC\x02\x00" + (bsize - 1).to_bytes(2, "little") + extra_after
        return hdr + z + zlib.crc32(piece).to_bytes(4, "little") + len(piece).to_bytes(4, "little")

    def index(img, pos=0, max_out=1 << 40, cap=64):
        tab = (_lib.BgzfBlock * cap)()
        nb, nxt, outb = C.c_int64(), C.c_int64(), C.c_int64()
        rc = L.exb_bgzf_index_host(img, len(img), pos, max_out, tab, cap, C.byref(nb), C.byref(nxt), C.byref(outb))
        return rc, [(tab[k].in_off, tab[k].out_off, tab[k].clen, tab[k].isize) for k in range(nb.value)], nxt.value, outb.value

    a, b, c = b"A" * 1000, b"CG" * 700, b"T" * 10
    other = b"XY\x03\x00abc"  # a foreign subfield: SI1 SI2 SLEN data
    img = member(a, extra_before=other) + member(b, extra_after=other) + member(c)
    assert L.exb_bgzf_probe_host(img, len(img)) == 1
    rc, tab, nxt, outb = index(img)
    assert rc == 0 and nxt == len(img) and outb == 2410 and [t[3] for t in tab] == [1000, 1400, 10] and [t[1] for t in tab] == [0, 1000, 2400]
    host = _host()
    for (in_off, _, clen, isize), want in zip(tab, (a, b, c)):
        out = C.create_string_buffer(isize + 8)
        assert host.ifl_host_inflate(img[in_off:in_off + clen], clen, out, isize, 0) == 0 and out.raw[:isize] == want
    # cuts: by text size (at least one member is always taken) and by table capacity; in_off is relative to `pos`
    rc, tab, nxt, outb = index(img, max_out=1500)
    assert rc == 0 and len(tab) == 1 and outb == 1000
    rc, tab2, nxt2, outb2 = index(img, pos=nxt, max_out=10)
    assert rc == 0 and len(tab2) == 1 and outb2 == 1400 and tab2[0][0] == 12 + 6 + len(other) and tab2[0][1] == 0
    rc, tab3, _, _ = index(img, cap=2)
    assert rc == 0 and len(tab3) == 2
    # not BGZF: plain gzip, a member with FNAME, no BC subfield; truncated; a header that lies about the text size
    import gzip
    assert L.exb_bgzf_probe_host(gzip.compress(a), 40) == 0
    assert L.exb_bgzf_probe_host(member(a, flags=4 | 8), 40) == 0
    assert L.exb_bgzf_probe_host(img[:10], 10) == 0
    assert index(img + gzip.compress(a))[0] != 0 and b"not a BGZF block header" in L.exb_last_error()
    assert index(img[:-5])[0] != 0 and b"truncated BGZF block" in L.exb_last_error()
    big = bytearray(member(a))
    big[-4:] = (70000).to_bytes(4, "little")
    assert index(bytes(big))[0] != 0 and b"claims 70000 bytes" in L.exb_last_error()


def test_bgzip_tool_rewrites_gzip_as_bgzf(tmp_path):
    """tools/bgzip.py: a single-stream .gz (which keeps the streaming host decoder) rewritten as BGZF (which the reader inflates
    on the device); host-side only."""
    import gzip
    from exon_duckdb_b200 import _lib
    from tools import bgzip
    fq, _ = util.random_fastq(47, 4000, min_len=50, max_len=150, tricky=False)
    src = tmp_path / "plain.fastq.gz"
    src.write_bytes(gzip.compress(fq))
    dst = tmp_path / "blocked.fastq.gz"
    assert _lib.lib().exb_bgzf_probe_host(src.read_bytes(), 64) == 0
    t, f = bgzip.recompress(str(src), str(dst), threads=3)
    img = dst.read_bytes()
    assert t == len(fq) and f == len(img)
    assert _lib.lib().exb_bgzf_probe_host(img, len(img)) == 1
    assert gzip.decompress(img) == fq
    assert img.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    raw = tmp_path / "raw.fasta"
    raw.write_bytes(b">a\nACGT\n")
    bgzip.recompress(str(raw), str(dst))
    assert gzip.decompress(dst.read_bytes()) == b">a\nACGT\n"
