"""The CPU oracle against the reference's own fixtures and known answers.

Every expectation here is copied from a reference test file (cited) or from the
outputs of the reference's unmodified scalar-function sources compiled against
the vendored DuckDB (tests/golden/scalar_ref_vectors.json, made by
oracle/make_scalar_vectors.py).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

SEQ60 = b"GATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT"
QUAL60 = b"!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65"


def _read(golden_dir, name):
    with open(os.path.join(golden_dir, name), "rb") as f:
        return f.read()


def test_fastq_fixture_rows(golden_dir):
    # test_fastq_scan.test:5-8 (count = 2) and :34-41 (first row, 4 columns in this order)
    t = O.parse_fastq(_read(golden_dir, "test.fastq"))
    assert t.n == 2
    assert t.names == ["name", "description", "sequence", "quality_scores"]
    rows = t.rows()
    assert rows[0] == (b"SEQ_ID", b"This is a description", SEQ60, QUAL60)
    assert rows[1] == (b"SEQ_ID2", None, SEQ60, QUAL60)


def test_fastq_plus_line_content_is_dropped(golden_dir):
    # test2.fastq carries "+This is a description" on the plus line; it is not a column
    t = O.parse_fastq(_read(golden_dir, "test2.fastq"))
    assert t.rows() == [(b"SEQ_ID", None, SEQ60, QUAL60), (b"SEQ_ID2", None, SEQ60, QUAL60)]


def test_fastq_directory_fixtures(golden_dir):
    # test_fastq_scan.test:64-68: the two files of fastq/ hold 4 records
    n = sum(O.parse_fastq(_read(golden_dir, "fastq/" + f)).n for f in ("copy-a.fastq", "copy-b.fastq"))
    assert n == 4


def test_fasta_fixture_rows(golden_dir):
    # test_fasta_scan.test:5-8 (count 2), :34-37 (column `id`, WHERE id = 'a' -> 1 row)
    t = O.parse_fasta(_read(golden_dir, "test.fasta"))
    assert t.names == ["id", "description", "sequence"]
    assert t.rows() == [(b"a", b"description", b"ATCG"), (b"b", b"description2", b"ATCG")]
    assert sum(1 for r in t.rows() if r[0] == b"a") == 1


def test_fasta_missing_description_is_null(golden_dir):
    # test_fasta_copy.test:74-80 documents `description IS NULL` for test.mixed-desc.fasta
    t = O.parse_fasta(_read(golden_dir, "test.mixed-desc.fasta"))
    assert t.rows() == [(b"a", b"description", b"ATCG"), (b"b", None, b"ATCG")]


def test_fasta_wrapped_and_crlf():
    t = O.parse_fasta(b">x d1 d2 \r\nACGT\r\nAC\r\n\r\n>y\nGG\n\nTT")
    assert t.rows() == [(b"x", b"d1 d2", b"ACGTAC"), (b"y", None, b"GGTT")]


def test_fastq_errors():
    with pytest.raises(O.OracleError):
        O.parse_fastq(b"SEQ\nACGT\n+\nIIII\n")  # no '@'
    with pytest.raises(O.OracleError):
        O.parse_fastq(b"@a\nACGT\n-\nIIII\n")  # no '+'
    with pytest.raises(O.OracleError):
        O.parse_fastq(b"@a\nACGT\n+\n")  # truncated
    assert O.parse_fastq(b"").n == 0
    assert O.parse_fastq(b"@a\nACGT\n+\nIIII").rows() == [(b"a", None, b"ACGT", b"IIII")]


# ---- scalar known answers (test_scalar_functions.test:5-46 and SURVEY 8c [VERIFIED-RUN])
@pytest.mark.parametrize("seq,want", [
    (b"ATGC", 0.5), (b"ATGCGC", np.float32(4) / np.float32(6)), (b"GGA", np.float32(2) / np.float32(3)),
    (b"gcGC", 0.5), (b"GCN", np.float32(2) / np.float32(3)), (b"ATGCGCA", np.float32(4) / np.float32(7)),
    (b"", 0.0), (b"GGGG", 1.0), (b"ATCG", 0.5),
])
def test_gc_content_known_answers(seq, want):
    got = O.gc_content(seq)
    assert got.dtype == np.float32
    assert got == np.float32(want)


def test_gc_content_null():
    assert O.gc_content(None) is None  # test_scalar_functions.test:17-21


def test_reverse_complement_known_answers():
    # :41-46 -- the reference's table is A->C T->G C->A G->T with no reversal (SURVEY finding 3)
    assert O.reverse_complement(b"ATCG") == b"CGAT"
    assert O.reverse_complement(b"GGGG") == b"TTTT"
    assert O.reverse_complement(b"ATGC") == b"CGTA"
    assert O.reverse_complement(b"AACG") == b"CCAT"
    assert O.reverse_complement(b"ATGCGC") == b"CGTATA"
    assert O.reverse_complement(b"") == b""
    for bad in (b"acgt", b"ACGN"):
        with pytest.raises(O.InvalidInput):
            O.reverse_complement(bad)


def test_complement_known_answers():
    # :30-39
    assert O.complement(b"ATGC") == b"TACG"
    assert O.complement(b"ATGCGC") == b"TACGCG"
    with pytest.raises(O.InvalidInput):
        O.complement(b"ATCGQ")


def test_quality_decode_known_answers():
    assert O.quality_score_string_to_list(b"!'*5I~").tolist() == [0, 6, 9, 20, 40, 93]
    assert O.mean_quality(b"IIII5555") == 30.0
    assert O.quality_score_string_to_list(bytes([0x80, 0xFF])).tolist() == [-128 - 33, -1 - 33]  # char is signed


def test_scalar_vectors_from_reference_build(golden_dir):
    """Vectors produced by the reference's own module.cpp files running inside DuckDB v0.8.1."""
    path = os.path.join(golden_dir, "scalar_ref_vectors.json")
    if not os.path.exists(path):
        pytest.skip("scalar_ref_vectors.json not generated")
    with open(path) as f:
        vec = json.load(f)
    assert len(vec["cases"]) > 100
    for case in vec["cases"]:
        s = case["seq"].encode("latin-1")
        if "gc" in case:
            assert O.gc_content(s) == np.float32(case["gc"]), s
        if "rc" in case:
            if case["rc"] is None:
                with pytest.raises(O.InvalidInput):
                    O.reverse_complement(s)
            else:
                assert O.reverse_complement(s) == case["rc"].encode("latin-1")
        if "comp" in case:
            if case["comp"] is None:
                with pytest.raises(O.InvalidInput):
                    O.complement(s)
            else:
                assert O.complement(s) == case["comp"].encode("latin-1")
        for key, fn in (("tr", O.transcribe), ("rtr", O.reverse_transcribe)):
            if key in case:
                if case[key] is None:
                    with pytest.raises(O.InvalidInput):
                        fn(s)
                else:
                    assert fn(s) == case[key].encode("latin-1")
        if "aa" in case:
            assert O.translate_dna_to_aa(s) == case["aa"].encode("latin-1")
        if "aa_error" in case:
            with pytest.raises(O.InvalidInput) as ei:
                O.translate_dna_to_aa(s)
            assert str(ei.value).encode("latin-1", "replace") == case["aa_error"].encode("latin-1") or str(ei.value) == case["aa_error"]
        if "qual" in case:
            assert O.quality_score_string_to_list(s).tolist() == case["qual"]
        if "mean_q" in case:
            assert O.mean_quality(s) == case["mean_q"]


# ------------------------------------------------------------------ writers (parity unpinned, see exon_oracle.c)
def test_oracle_writers_are_the_inverse_of_the_parsers(golden_dir):
    import exb_testutil as util

    # the reference's canonical fixtures come back byte for byte
    text = open(os.path.join(golden_dir, "test.fastq"), "rb").read()
    t = O.parse_fastq(text)
    assert O.format_fastq(t.strings("name"), t.strings("description"), t.strings("sequence"), t.strings("quality_scores")) == text
    for name in ("test.fasta", "test.mixed-desc.fasta"):
        text = open(os.path.join(golden_dir, name), "rb").read()
        t = O.parse_fasta(text)
        img = O.format_fasta(t.strings("id"), t.strings("description"), t.strings("sequence"))
        assert O.parse_fasta(img).rows() == t.rows()
    # random records: parse(format(x)) == x (a NULL and an empty description are the same on disk: both parse as NULL)
    _, recs = util.random_fastq(5, 500, tricky=False)
    img = O.format_fastq([r[0] for r in recs], [r[1] for r in recs], [r[2] for r in recs], [r[3] for r in recs])
    back = O.parse_fastq(img)
    assert back.rows() == [(r[0], r[1] if r[1] else None, r[2], r[3]) for r in recs]
    # FASTA wrapping: every sequence line but the last holds exactly line_width bases
    seqs = [b"A" * n for n in (0, 1, 59, 60, 61, 120, 121)]
    img = O.format_fasta([b"s%d" % i for i in range(len(seqs))], [None] * len(seqs), seqs, line_width=60)
    assert img == b"".join(b">s%d\n" % i + b"".join(s[k:k + 60] + b"\n" for k in range(0, len(s), 60)) for i, s in enumerate(seqs))
    assert O.parse_fasta(img).strings("sequence") == seqs


def test_writer_without_a_gpu_is_a_loud_error(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    import ctypes as C
    from exon_duckdb_b200 import _lib
    w = C.c_void_p()
    rc = _lib.lib().exb_writer_open(str(tmp_path / "x.fastq").encode(), b"fastq", None, 0, 0, C.byref(w))
    assert rc == _lib.ERR_CUDA and b"no CUDA device" in _lib.lib().exb_last_error()
