"""The DuckDB-facing boundary: `LOAD exon`, read_fasta / read_fastq, replacement scans and the scalar functions.

Two extensions are driven through build/rt/sqlrun (a DuckDB v0.8.1 host, tools/sqlrun.cpp):
  PRODUCT  exon_duckdb_b200/duckdb_ext/exon.duckdb_extension  -- our host code over the exb_reader_* C ABI;
  REFGLUE  oracle/_ref/exon.duckdb_extension                   -- the REFERENCE's unmodified C++ glue
           (arrow_table_function/module.cpp) linked against libexon_b200.so's new_reader / replacement_scan:
           the drop-in proof.  Its scalar functions are the reference's own CPU code (the live oracle).
The queries and expectations replay the reference's sqllogictests
(test/sql/exondb-release-with-deb-info/test_fastq_scan.test, test_fasta_scan.test, test_scalar_functions.test);
zstd cases included (the system's libzstd.so.1 is bound at run time).

The binaries are built in the development container (they need the reference's vendored DuckDB headers) and travel
to the GPU box with the snapshot; the tests skip when they are absent.
"""
import json
import os
import random
import subprocess

import numpy as np
import pytest

import exb_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SQLRUN = os.path.join(ROOT, "build", "rt", "sqlrun")
PRODUCT = os.path.join(ROOT, "exon_duckdb_b200", "duckdb_ext", "exon.duckdb_extension")
REFGLUE = os.path.join(ROOT, "oracle", "_ref", "exon.duckdb_extension")
G = os.path.join(ROOT, "tests", "golden")

SEQ60 = "GATTTGGGGTTCAAAGCAGTATCGATCAAATAGTAAATCCATTTGTTCAACTCACAGTTT"
QUAL60 = "!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65"


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip("%s not built (needs the reference's DuckDB headers: bash oracle/build_ref.sh; make -C exon_duckdb_b200/duckdb_ext)" % os.path.relpath(p, ROOT))


def run_sql(ext, statements, threads=None, env=None):
    _need(SQLRUN, ext)
    text = "LOAD '%s';\n" % ext + "\n".join(s.rstrip().rstrip(";") + ";" for s in statements) + "\n"
    cmd = [SQLRUN] + (["-threads", str(threads)] if threads else [])
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run(cmd, input=text.encode("utf-8"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, env=e)
    assert out.returncode == 0, out.stderr.decode()
    res = [json.loads(l) for l in out.stdout.decode("utf-8").splitlines()]
    assert res and res[0]["ok"], res[:1]
    assert len(res) == len(statements) + 1, (len(res), len(statements))
    return res[1:]


def rows(r):
    assert r["ok"], r.get("error")
    return r["rows"]


def scalar(r):
    return rows(r)[0][0]


# ------------------------------------------------------------------ no GPU needed: registration, bind, plans
def test_load_registers_the_path_functions():
    r = run_sql(PRODUCT, ["SELECT function_name, function_type, return_type FROM duckdb_functions() WHERE function_name IN "
                          "('read_fasta','read_fastq','gc_content','reverse_complement','complement','quality_score_string_to_list') ORDER BY 1"])
    assert rows(r[0]) == [["complement", "scalar", "VARCHAR"], ["gc_content", "scalar", "FLOAT"],
                          ["quality_score_string_to_list", "scalar", "INTEGER[]"], ["read_fasta", "table", None],
                          ["read_fastq", "table", None], ["reverse_complement", "scalar", "VARCHAR"]]


def test_gpu_stats_table_function_is_registered():
    """SURVEY 5 'Metrics / logging': exon_gpu_stats() -- per-scan counters; empty before any scan ran."""
    r = run_sql(PRODUCT, ["DESCRIBE SELECT * FROM exon_gpu_stats()", "SELECT count(*) FROM exon_gpu_stats()"])
    cols = [(x[0], x[1]) for x in rows(r[0])]
    assert cols[:3] == [("path", "VARCHAR"), ("files", "INTEGER"), ("format", "VARCHAR")]
    assert ("rows", "BIGINT") in cols and ("io_path", "VARCHAR") in cols and ("seconds_scan", "DOUBLE") in cols and ("gb_per_s", "DOUBLE") in cols
    assert rows(r[1]) == [["0"]]


@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_schema_matches_the_reference(ext):
    # FileTypeBind: FASTQ (name, description, sequence, quality_scores), FASTA (id, description, sequence), all VARCHAR
    r = run_sql(ext, ["DESCRIBE SELECT * FROM read_fastq('%s/test.fastq')" % G, "DESCRIBE SELECT * FROM '%s/test.fasta'" % G,
                      "DESCRIBE SELECT * FROM read_fasta('%s/test.fasta.gzip', compression='gzip')" % G])
    assert [(x[0], x[1]) for x in rows(r[0])] == [("name", "VARCHAR"), ("description", "VARCHAR"), ("sequence", "VARCHAR"), ("quality_scores", "VARCHAR")]
    assert [(x[0], x[1]) for x in rows(r[1])] == [("id", "VARCHAR"), ("description", "VARCHAR"), ("sequence", "VARCHAR")]
    assert [(x[0], x[1]) for x in rows(r[2])] == [("id", "VARCHAR"), ("description", "VARCHAR"), ("sequence", "VARCHAR")]


@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_missing_file_is_a_bind_error(ext):
    # test_fastq_scan.test:61-62, test_fasta_scan.test:51-53: `statement error`
    r = run_sql(ext, ["SELECT count(*) FROM read_fastq('')", "SELECT count(*) FROM read_fasta('')", "SELECT count(*) FROM read_fastq('/nonexistent/x.fastq')"])
    assert not r[0]["ok"] and not r[1]["ok"] and not r[2]["ok"]


def test_complex_filters_are_absorbed_by_the_scan():
    r = run_sql(PRODUCT, [
        "EXPLAIN SELECT count(*) FROM read_fastq('%s/test.fastq') WHERE list_avg(quality_score_string_to_list(quality_scores)) > 30" % G,
        "EXPLAIN SELECT name FROM read_fastq('%s/test.fastq') WHERE gc_content(sequence) >= 0.25 AND name = 'SEQ_ID'" % G,
        "EXPLAIN SELECT name FROM read_fastq('%s/test.fastq') WHERE length(sequence) > 3" % G])
    plan0, plan1, plan2 = (scalar(x) if len(rows(x)[0]) == 1 else rows(x)[0][1] for x in r)
    assert "Device filters" in plan0 and "mean_quality" in plan0 and "FILTER" not in plan0
    assert "gc_content" in plan1 and "Device filters" in plan1 and "FILTER" not in plan1
    assert "Device filters" not in plan2 and "FILTER" in plan2  # length() counts code points, not bytes: stays in DuckDB


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = run_sql(PRODUCT, ["SELECT count(*) FROM read_fastq('%s/test.fastq')" % G, "SELECT gc_content('ATGC')"])
    assert not r[0]["ok"] and "no CUDA device" in r[0]["error"]
    assert not r[1]["ok"] and "no CUDA device" in r[1]["error"]


def test_copy_functions_bind_like_the_removed_reference_writers(tmp_path):
    # test_fasta_copy.test:43-50 (commented out in the reference): an existing target is an error unless FORCE is given;
    # unknown options, wrong column counts and non-VARCHAR columns are bind errors.  No GPU is touched at bind.
    existing = tmp_path / "there.fasta"
    existing.write_bytes(b">a\nACGT\n")
    r = run_sql(PRODUCT, [
        "COPY (SELECT * FROM read_fasta('%s/test.fasta')) TO '%s' (FORMAT 'fasta')" % (G, existing),
        "COPY (SELECT 1 AS a, 'x' AS b, 'y' AS c) TO '%s/n.fasta' (FORMAT 'fasta')" % tmp_path,
        "COPY (SELECT 'a' AS a, 'x' AS b) TO '%s/n.fastq' (FORMAT 'fastq')" % tmp_path,
        "COPY (SELECT 'a' AS a, 'x' AS b, 'y' AS c) TO '%s/n.fasta' (FORMAT 'fasta', NOPE 1)" % tmp_path,
        "COPY (SELECT 'a' AS a, 'x' AS b, 'y' AS c) TO '%s/n.fasta' (FORMAT 'fasta', COMPRESSION 'brotli')" % tmp_path,
    ])
    assert not r[0]["ok"] and "exists" in r[0]["error"] and "FORCE" in r[0]["error"]
    assert not r[1]["ok"] and "VARCHAR" in r[1]["error"]
    assert not r[2]["ok"] and "4 columns" in r[2]["error"]
    assert not r[3]["ok"] and "Unknown option" in r[3]["error"]
    assert not r[4]["ok"]
    assert existing.read_bytes() == b">a\nACGT\n"


# ------------------------------------------------------------------ GPU: the reference's sqllogictests replayed
FASTQ_SCAN = [  # test_fastq_scan.test
    ("SELECT count(*) FROM read_fastq('{G}/test.fastq')", [["2"]]),
    ("SELECT count(*) FROM read_fastq('{G}/test.fastq.gz')", [["2"]]),
    ("SELECT count(*) FROM read_fastq('{G}/test.fastq.gzip', compression='gzip')", [["2"]]),
    ("SELECT count(*) FROM read_fastq('{G}/test.fastq.zst')", [["2"]]),                          # :22-26
    ("SELECT count(*) FROM read_fastq('{G}/test.fastq.zstd', compression='zstd')", [["2"]]),     # :28-32
    ("SELECT count(*) FROM '{G}/test.fastq.zst'", [["2"]]),                                      # :55-59
    ("SELECT * FROM read_fastq('{G}/test.fastq') LIMIT 1", [["SEQ_ID", "This is a description", SEQ60, QUAL60]]),
    ("SELECT count(*) FROM '{G}/test.fastq'", [["2"]]),
    ("SELECT count(*) FROM '{G}/test.fastq.gz'", [["2"]]),
    ("SELECT COUNT(*) FROM read_fastq('{G}/fastq/') LIMIT 1", [["4"]]),
]
FASTA_SCAN = [  # test_fasta_scan.test
    ("SELECT count(*) FROM read_fasta('{G}/test.fasta')", [["2"]]),
    ("SELECT count(*) FROM read_fasta('{G}/test.fasta.gzip', compression='gzip')", [["2"]]),
    ("SELECT count(*) FROM read_fasta('{G}/test.fasta.gz')", [["2"]]),
    ("SELECT count(*) FROM read_fasta('{G}/test.fasta.zstd', compression='zstd')", [["2"]]),     # :22-26
    ("SELECT count(*) FROM read_fasta('{G}/test.fasta.zst')", [["2"]]),                          # :45-49
    ("SELECT count(*) FROM '{G}/test.fasta'", [["2"]]),
    ("SELECT count(*) FROM '{G}/test.fasta' WHERE id = 'a'", [["1"]]),
    ("SELECT count(*) FROM '{G}/test.fasta.gz'", [["2"]]),
    ("SELECT COUNT(*) FROM read_fasta('{G}/fasta/', compression='gzip')", [["4"]]),
    # test_fasta_copy.test:74-80 (documented expectation): a header without description gives NULL
    ("SELECT id, description, sequence FROM read_fasta('{G}/test.mixed-desc.fasta')", [["a", "description", "ATCG"], ["b", None, "ATCG"]]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_reference_scan_sqllogictests(cuda_device, ext):
    cases = FASTQ_SCAN + FASTA_SCAN
    res = run_sql(ext, [q.format(G=G) for q, _ in cases])
    for (q, want), r in zip(cases, res):
        assert rows(r) == want, q


@pytest.mark.gpu
def test_reference_scalar_sqllogictests(cuda_device):
    # test_scalar_functions.test:5-46 through the PRODUCT's CUDA scalar functions
    r = run_sql(PRODUCT, [
        "SELECT gc_content(seq) FROM (SELECT 'ATGC' AS seq UNION ALL SELECT 'ATGCGC' AS seq)",
        "SELECT gc_content('')", "SELECT gc_content(NULL) IS NULL",
        "WITH two_seqs AS (SELECT 'ATCG' AS sequence UNION ALL SELECT 'GGGG') SELECT gc_content(sequence) FROM two_seqs",
        "SELECT complement(seq) FROM (SELECT 'ATGC' AS seq UNION ALL SELECT 'ATGCGC' AS seq)",
        "SELECT complement('ATCGQ')",
        "SELECT reverse_complement(seq) FROM (SELECT 'ATCG' AS seq UNION ALL SELECT 'GGGG' AS seq)",
        "SELECT quality_score_string_to_list('!''*5I~')",
        "SELECT list_avg(quality_score_string_to_list('IIII5555'))",
        "SELECT typeof(quality_score_string_to_list('II'))",
    ])
    assert rows(r[0]) == [["0.5"], ["0.6666667"]]
    assert rows(r[1]) == [["0.0"]] and rows(r[2]) == [["true"]]
    assert rows(r[3]) == [["0.5"], ["1.0"]]
    assert rows(r[4]) == [["TACG"], ["TACGCG"]]
    assert not r[5]["ok"] and "Invalid character in sequence: Q" in r[5]["error"]
    assert rows(r[6]) == [["CGAT"], ["TTTT"]]
    assert rows(r[7]) == [["[0, 6, 9, 20, 40, 93]"]] and rows(r[8]) == [["30.0"]] and rows(r[9]) == [["INTEGER[]"]]


@pytest.mark.gpu
def test_reference_scalar_sqllogictests_transcribe_translate(cuda_device):
    # test_scalar_functions.test:48-89 through the PRODUCT's CUDA scalar functions
    allc = ("AAAAATAACAAGATAATTATCATGACAACTACCACGAGAAGTAGCAGGTAATATTACTAGTTATTTTTCTTGTCATCTTCCTCGTGATGTTGCTGGCAACATCACCAGC"
            "TACTTCTCCTGCCACCTCCCCCGCGACGTCGCCGGGAAGATGACGAGGTAGTTGTCGTGGCAGCTGCCGCGGGAGGTGGCGGG")
    r = run_sql(PRODUCT, [
        "SELECT transcribe(t) FROM (SELECT 'ATCG' AS t UNION ALL SELECT 'ATCGATCG' AS t)",
        "SELECT transcribe('ATNN')",
        "SELECT translate_dna_to_aa(seq) FROM (SELECT 'ATGCGC' AS seq UNION ALL SELECT 'ATGCGC' AS seq)",
        "SELECT translate_dna_to_aa('%s')" % allc,
        "SELECT translate_dna_to_aa('NNN')",
        "SELECT translate_dna_to_aa('ATTT')",
        "SELECT reverse_transcribe(seq) FROM (SELECT 'AUCG' AS seq UNION ALL SELECT 'AUCU' AS seq)",
        "SELECT reverse_transcribe('AUNN')",
        "SELECT translate_dna_to_aa(NULL) IS NULL, translate_dna_to_aa('')",
    ])
    assert rows(r[0]) == [["AUCG"], ["AUCGAUCG"]]
    assert not r[1]["ok"] and "Invalid character in sequence: N" in r[1]["error"]
    assert rows(r[2]) == [["MR"], ["MR"]]
    assert rows(r[3]) == [["KNNKIIIMTTTTRSSR*YY*LFFLSSSS*CCWQHHQLLLLPPPPRRRREDDEVVVVAAAAGGGG"]]
    assert not r[4]["ok"] and "Invalid codon: NNN" in r[4]["error"]
    assert not r[5]["ok"] and "Invalid sequence length: 4" in r[5]["error"]
    assert rows(r[6]) == [["ATCG"], ["ATCT"]]
    assert not r[7]["ok"] and "Invalid character in sequence: N" in r[7]["error"]
    assert rows(r[8]) == [["true", ""]]


@pytest.mark.gpu
def test_transcribe_translate_against_reference_vectors(cuda_device):
    """transcribe / reverse_transcribe / translate_dna_to_aa of the PRODUCT inside DuckDB vs the reference's own functions
    (tests/golden/scalar_ref_vectors.json), the valid inputs in ONE multi-row table, the invalid ones one by one."""
    with open(os.path.join(G, "scalar_ref_vectors.json")) as f:
        cases = [c for c in json.load(f)["cases"] if "tr" in c and len(c["seq"]) < 6000]
    lit = lambda s: "'" + s.encode("latin-1").decode("utf-8").replace("'", "''") + "'"
    for key, fn in (("tr", "transcribe"), ("rtr", "reverse_transcribe"), ("aa", "translate_dna_to_aa")):
        good = [c for c in cases if c.get(key) is not None]
        assert len(good) > 20, key
        values = ", ".join("(%d, %s)" % (i, lit(c["seq"])) for i, c in enumerate(good))
        r = run_sql(PRODUCT, ["CREATE TABLE s AS SELECT * FROM (VALUES %s) t(i, seq)" % values, "SELECT i, %s(seq) FROM s ORDER BY i" % fn])
        got = rows(r[1])
        assert len(got) == len(good)
        for (i, v), c in zip(got, good):
            assert v.encode("utf-8").decode("latin-1") == c[key], (fn, c["seq"][:40])
    bad = [c for c in cases if "aa_error" in c][:12]
    r = run_sql(PRODUCT, ["SELECT translate_dna_to_aa(%s)" % lit(c["seq"]) for c in bad])
    for x, c in zip(r, bad):
        assert not x["ok"] and c["aa_error"].encode("latin-1").decode("utf-8") in x["error"], (c["seq"][:40], x)


@pytest.mark.gpu
def test_scalar_functions_against_reference_vectors(cuda_device):
    """The product's CUDA scalar functions inside DuckDB vs the reference's own functions (golden vectors), in ONE multi-row table."""
    with open(os.path.join(G, "scalar_ref_vectors.json")) as f:
        cases = json.load(f)["cases"]
    seqs = [c for c in cases if "gc" in c and len(c["seq"]) < 6000]
    lit = lambda s: "'" + s.encode("latin-1").decode("utf-8").replace("'", "''") + "'"
    values = ", ".join("(%d, %s)" % (i, lit(c["seq"])) for i, c in enumerate(seqs))
    ok = [i for i, c in enumerate(seqs) if c.get("rc") is not None]
    quals = [c for c in cases if "qual" in c]
    qvalues = ", ".join("(%d, %s)" % (i, lit(c["seq"])) for i, c in enumerate(quals))
    r = run_sql(PRODUCT, [
        "CREATE TABLE s AS SELECT * FROM (VALUES %s) t(i, seq)" % values,
        "SELECT i, gc_content(seq)::DOUBLE FROM s ORDER BY i",
        "SELECT i, reverse_complement(seq), complement(seq) FROM s WHERE i IN (%s) ORDER BY i" % ",".join(map(str, ok)),
        "CREATE TABLE q AS SELECT * FROM (VALUES %s) t(i, qual)" % qvalues,
        "SELECT i, quality_score_string_to_list(qual), list_avg(quality_score_string_to_list(qual)) FROM q ORDER BY i",
    ])
    got = rows(r[1])
    assert len(got) == len(seqs)
    for (i, g), c in zip(got, seqs):
        assert np.float32(float(g)) == np.float32(c["gc"]), c["seq"][:40]
    for (i, rc, comp) in rows(r[2]):
        c = seqs[int(i)]
        assert rc.encode("utf-8").decode("latin-1") == c["rc"] and comp.encode("utf-8").decode("latin-1") == c["comp"]
    for (i, lst, avg), c in zip(rows(r[4]), quals):
        assert json.loads(lst) == c["qual"]
        assert float(avg) == c["mean_q"]
    bad = [c for c in seqs if c.get("rc") is None][:5]
    r = run_sql(PRODUCT, ["SELECT reverse_complement(%s)" % lit(c["seq"]) for c in bad])
    assert all((not x["ok"]) and "Invalid character in sequence" in x["error"] for x in r)


def _write(tmp_path, name, data):
    p = os.path.join(str(tmp_path), name)
    with open(p, "wb") as f:
        f.write(data)
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_file_order_survives_order_preserving_sinks(cuda_device, tmp_path, ext):
    """get_batch_index / cardinality are registered like the reference's (module.cpp:307-308): with several DuckDB threads,
    CREATE TABLE AS, LIMIT / OFFSET and rowid see the records in file order."""
    from oracle import oracle as O
    text, _ = util.random_fastq(17, 9000, min_len=1, max_len=120, tricky=False)
    path = _write(tmp_path, "order.fastq", text)
    names = [x.decode("latin-1") for x in O.parse_fastq(text).strings("name")]
    r = run_sql(ext, [
        "CREATE TABLE t AS SELECT name FROM read_fastq('%s')" % path,
        "SELECT name FROM t WHERE rowid IN (0, 2047, 2048, 4999, 8999) ORDER BY rowid",
        "SELECT name FROM read_fastq('%s') LIMIT 3 OFFSET 6000" % path,
        "SELECT count(*) FROM t",
    ], threads=4, env={"EXON_B200_CHUNK_BYTES": str(100_000)})
    assert rows(r[1]) == [[names[i]] for i in (0, 2047, 2048, 4999, 8999)]
    assert rows(r[2]) == [[n] for n in names[6000:6003]]
    assert rows(r[3]) == [["9000"]]


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_fastq_rows_filters_and_projection(cuda_device, tmp_path, ext):
    """A multi-chunk synthetic file through DuckDB: every row, projections, simple + complex filters, COUNT(*), vs the oracle."""
    from oracle import oracle as O
    text, _ = util.random_fastq(11, 5000, min_len=1, max_len=200, tricky=False)
    path = _write(tmp_path, "r.fastq", text)
    ref = O.parse_fastq(text)
    want = [tuple(None if v is None else v.decode() for v in row) for row in ref.rows()]
    env = {"EXON_B200_CHUNK_BYTES": str(200_000)}  # several device chunks, records straddling their edges
    mq = [O.mean_quality_pass(q, ">", 60.0) for q in ref.strings("quality_scores")]
    gcv = [O.gc_content(s) for s in ref.strings("sequence")]
    res = run_sql(ext, [
        "SELECT count(*) FROM read_fastq('%s')" % path,
        "SELECT * FROM read_fastq('%s')" % path,
        "SELECT sequence, name FROM read_fastq('%s') WHERE description IS NULL" % path,
        "SELECT count(*) FROM read_fastq('%s') WHERE list_avg(quality_score_string_to_list(quality_scores)) > 60" % path,
        "SELECT name FROM read_fastq('%s') WHERE list_avg(quality_score_string_to_list(quality_scores)) > 60" % path,
        # the reference's own gc_content collapses multi-row chunks (SURVEY finding 4): only the product evaluates this one per row
        ("SELECT count(*), sum(length(sequence)) FROM read_fastq('%s') WHERE gc_content(sequence) < 0.45" if ext == PRODUCT else "SELECT 0, 0 FROM read_fastq('%s') LIMIT 0") % path,
    ], env=env)
    assert rows(res[0]) == [[str(len(want))]]
    assert [tuple(r) for r in rows(res[1])] == want
    assert [tuple(r) for r in rows(res[2])] == [(w[2], w[0]) for w in want if w[1] is None]
    assert rows(res[3]) == [[str(sum(mq))]]
    assert [r[0] for r in rows(res[4])] == [w[0] for w, m in zip(want, mq) if m]
    # DuckDB compares FLOAT with the constant cast to FLOAT (checked: 0.45::FLOAT < 0.45 is false)
    sel = [w for w, g in zip(want, gcv) if g < np.float32(0.45)]
    if ext == PRODUCT:
        assert rows(res[5]) == [[str(len(sel)), str(sum(len(w[2]) for w in sel))]]


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_fasta_rows_and_gc_per_contig(cuda_device, tmp_path, ext):
    from oracle import oracle as O
    text, _ = util.random_fasta(5, 300, min_len=0, max_len=5000, tricky=False)
    path = _write(tmp_path, "g.fasta", text)
    ref = O.parse_fasta(text)
    want = [tuple(None if v is None else v.decode() for v in row) for row in ref.rows()]
    env = {"EXON_B200_CHUNK_BYTES": str(150_000)}
    res = run_sql(ext, ["SELECT * FROM read_fasta('%s')" % path, "SELECT count(*) FROM read_fasta('%s')" % path,
                        "SELECT id FROM read_fasta('%s') WHERE id >= 's2' AND id < 's3'" % path], env=env)
    assert [tuple(r) for r in rows(res[0])] == want
    assert rows(res[1]) == [[str(len(want))]]
    assert [r[0] for r in rows(res[2])] == [w[0] for w in want if "s2" <= w[0] < "s3"]
    if ext == PRODUCT:  # C3's query; one query per row would pin the reference (its gc_content collapses chunks, SURVEY finding 4)
        r = run_sql(ext, ["SELECT id, gc_content(sequence)::DOUBLE FROM read_fasta('%s')" % path], env=env)
        for (i, g), w in zip(rows(r[0]), want):
            assert i == w[0] and np.float32(float(g)) == O.gc_content(w[2].encode())


# ------------------------------------------------------------------ projections served by the scan (optimizer extension)
def _plan(r):
    return scalar(r) if len(rows(r)[0]) == 1 else rows(r)[0][1]


def test_projections_are_fused_into_the_scan():
    """gc_content / the LUT maps / quality decode / list_avg(decode) over a scan column become computed columns of the scan
    (EXPLAIN lists them under "Device columns"); anything else keeps the scalar-function path."""
    fq, fa = "%s/test.fastq" % G, "%s/test.fasta" % G
    r = run_sql(PRODUCT, [
        "EXPLAIN SELECT id, gc_content(sequence) FROM read_fasta('%s')" % fa,
        "EXPLAIN SELECT avg(gc_content(sequence)) FROM read_fastq('%s')" % fq,
        "EXPLAIN SELECT sum(length(reverse_complement(sequence))) FROM read_fastq('%s')" % fq,
        "EXPLAIN SELECT name, list_avg(quality_score_string_to_list(quality_scores)) FROM read_fastq('%s')" % fq,
        # a LUT map raises on a byte outside its table: above a filter that DuckDB evaluates it must only see the surviving rows
        "EXPLAIN SELECT complement(sequence) FROM read_fastq('%s') WHERE length(sequence) > 3" % fq,
        # gc_content never raises: fused through the filter
        "EXPLAIN SELECT gc_content(sequence) FROM read_fastq('%s') WHERE length(sequence) > 3" % fq,
        # not a plain column underneath: the scalar function runs
        "EXPLAIN SELECT gc_content(sequence || 'A') FROM read_fastq('%s')" % fq,
        # DuckDB's common-subexpression pass decodes once and averages the list itself: the decode is the fused column
        "EXPLAIN SELECT quality_score_string_to_list(quality_scores), list_avg(quality_score_string_to_list(quality_scores)) FROM read_fastq('%s')" % fq,
    ])
    p = ["".join(ch for ch in _plan(x) if ch not in "│ \n─┌┐└┘┬┴") for x in r]  # the box renderer wraps long lines
    assert "Devicecolumns:gc_content(sequence)" in p[0]
    assert "Devicecolumns:gc_content(sequence)" in p[1] and "Notreadback:sequence" in p[1]
    assert "Devicecolumns:reverse_complement(se" in p[2] and "Notreadback:sequence" in p[2]
    assert "Devicecolumns:" in p[3] and "list_avg(quality_score_string_to_list(quality_scores))" in p[3] and "Notreadback:quality_scores" in p[3]
    assert "Devicecolumns" not in p[4] and "complement" in p[4]
    assert "Devicecolumns:gc_content(sequence)" in p[5] and "FILTER" in p[5] and "Notreadback" not in p[5]
    assert "Devicecolumns" not in p[6]
    assert "Devicecolumns:quality_" in p[7] and "Notreadback:quality_scores" in p[7] and "list_aggr(#" in p[7]


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 4])
def test_fused_projections_match_the_oracle(cuda_device, tmp_path, threads):
    """Every fused projection, row by row, against the oracle's scalar functions; with pushed-down filters, filters DuckDB
    keeps, several device chunks and several DuckDB threads."""
    from oracle import oracle as O
    text, _ = util.random_fastq(23, 6000, min_len=0, max_len=180, tricky=False)
    path = _write(tmp_path, "p.fastq", text)
    ref = O.parse_fastq(text)
    names = [x.decode() for x in ref.strings("name")]
    seqs, quals = ref.strings("sequence"), ref.strings("quality_scores")
    env = {"EXON_B200_CHUNK_BYTES": str(150_000)}
    res = run_sql(PRODUCT, [
        "SELECT name, gc_content(sequence)::DOUBLE, reverse_complement(sequence), complement(sequence) FROM read_fastq('%s')" % path,
        "SELECT name, quality_score_string_to_list(quality_scores), list_avg(quality_score_string_to_list(quality_scores)) FROM read_fastq('%s')" % path,
        "SELECT name, transcribe(sequence), gc_content(sequence)::DOUBLE FROM read_fastq('%s') WHERE list_avg(quality_score_string_to_list(quality_scores)) > 80" % path,
        "SELECT name, gc_content(sequence)::DOUBLE FROM read_fastq('%s') WHERE length(sequence) > 90" % path,
        "SELECT count(*), sum(length(reverse_complement(sequence))), avg(gc_content(sequence)) FROM read_fastq('%s')" % path,
        "SELECT name, gc_content(sequence)::DOUBLE, sequence FROM read_fastq('%s') WHERE name >= 'r3' LIMIT 5 OFFSET 2100" % path,
    ], threads=threads, env=env)
    got = rows(res[0])
    assert [g[0] for g in got] == names
    for g, s in zip(got, seqs):
        assert np.float32(float(g[1])) == O.gc_content(s) and g[2].encode() == O.reverse_complement(s) and g[3].encode() == O.complement(s)
    got = rows(res[1])
    assert [g[0] for g in got] == names
    for g, q in zip(got, quals):
        want = [c - 33 for c in q]
        assert json.loads(g[1]) == want if isinstance(g[1], str) else list(g[1]) == want
        if want:
            assert float(g[2]) == float(np.longdouble(sum(want)) / np.longdouble(len(want)))
        else:
            assert g[2] is None
    keep = [i for i, q in enumerate(quals) if O.mean_quality_pass(q, ">", 80.0)]
    got = rows(res[2])
    assert [g[0] for g in got] == [names[i] for i in keep]
    for g, i in zip(got, keep):
        assert g[1].encode() == seqs[i].replace(b"T", b"U") and np.float32(float(g[2])) == O.gc_content(seqs[i])
    keep = [i for i, s in enumerate(seqs) if len(s) > 90]
    got = rows(res[3])
    assert [g[0] for g in got] == [names[i] for i in keep]
    assert all(np.float32(float(g[1])) == O.gc_content(seqs[i]) for g, i in zip(got, keep))
    cnt, total, avg = rows(res[4])[0]
    assert (int(cnt), int(total)) == (len(seqs), sum(len(s) for s in seqs))
    assert abs(float(avg) - float(np.mean([np.float64(O.gc_content(s)) for s in seqs]))) < 1e-9
    keep = [i for i, n in enumerate(names) if n >= "r3"][2100:2105]
    assert [(g[0], g[2]) for g in rows(res[5])] == [(names[i], seqs[i].decode()) for i in keep]


@pytest.mark.gpu
def test_fused_map_reports_the_reference_error(cuda_device, tmp_path):
    """reverse_complement over the scan raises the reference's message for a byte outside ACGT (module.cpp:58-62), fused or not."""
    text = util.fastq_text([(b"a", None, b"ACGT", b"IIII"), (b"b", b"d", b"ACNT", b"IIII"), (b"c", None, b"ACGT", b"IIII")])
    path = _write(tmp_path, "n.fastq", text)
    _need(SQLRUN, PRODUCT)
    for q in ("SELECT reverse_complement(sequence) FROM read_fastq('%s')", "SELECT reverse_complement(sequence || '') FROM read_fastq('%s')"):
        out = subprocess.run([SQLRUN], input=("LOAD '%s';\n%s;\n" % (PRODUCT, q % path)).encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
        res = [json.loads(l) for l in out.stdout.decode().splitlines()]
        assert not res[1]["ok"] and "Invalid character in sequence: N" in res[1]["error"], res[1]
    r = run_sql(PRODUCT, ["SELECT reverse_complement(sequence) FROM read_fastq('%s') WHERE name <> 'b'" % path])
    assert rows(r[0]) == [["CATG"], ["CATG"]]  # the pushed-down filter removes the bad row before the map sees it (A->C C->A G->T T->G)


@pytest.mark.gpu
def test_fused_fasta_projections(cuda_device, tmp_path):
    from oracle import oracle as O
    text, _ = util.random_fasta(9, 200, min_len=0, max_len=4000, tricky=False)
    path = _write(tmp_path, "p.fasta", text)
    ref = O.parse_fasta(text)
    ids = [x.decode() for x in ref.strings("id")]
    seqs = ref.strings("sequence")
    res = run_sql(PRODUCT, ["SELECT id, gc_content(sequence)::DOUBLE FROM read_fasta('%s')" % path,
                            "SELECT id, gc_content(sequence)::DOUBLE, length(sequence) FROM read_fasta('%s') WHERE gc_content(sequence) > 0.5" % path,
                            "SELECT count(*), avg(gc_content(sequence)) FROM read_fasta('%s')" % path],
                  threads=4, env={"EXON_B200_CHUNK_BYTES": str(120_000)})
    assert [(g[0], np.float32(float(g[1]))) for g in rows(res[0])] == [(i, O.gc_content(s)) for i, s in zip(ids, seqs)]
    keep = [k for k, s in enumerate(seqs) if O.gc_content(s) > np.float32(0.5)]
    assert [(g[0], int(g[2])) for g in rows(res[1])] == [(ids[k], len(seqs[k])) for k in keep]
    assert int(rows(res[2])[0][0]) == len(seqs)


@pytest.mark.gpu
def test_c1_small_fasta_count(cuda_device, tmp_path):
    """BASELINE configs[0]: SELECT COUNT(*) FROM read_fasta on the 10 000-record synthetic FASTA, at its stated size."""
    from oracle import oracle as O
    from tools import synth
    p = synth.gen_params("fasta", 10_000, seed=1, len_min=200, len_max=2000, wrap=60)
    text = synth.gen_host(p).tobytes()
    path = _write(tmp_path, "c1.fasta", text)
    res = run_sql(PRODUCT, ["SELECT COUNT(*) FROM read_fasta('%s')" % path, "SELECT COUNT(*), SUM(length(sequence)) FROM '%s'" % path])
    ref = O.parse_fasta(text)
    assert rows(res[0]) == [["10000"]] and ref.n == 10_000
    assert rows(res[1]) == [["10000", str(sum(len(s) for s in ref.strings("sequence")))]]


@pytest.mark.gpu
def test_gpus_parameter(cuda_device, tmp_path):
    """`gpus := n` pins the number of device pipelines; more than the box has is a bind-time error, not a silent fallback."""
    import torch
    text, _ = util.random_fastq(31, 4000, min_len=20, max_len=150, tricky=False)
    path = _write(tmp_path, "g.fastq", text)
    n_dev = torch.cuda.device_count()
    r = run_sql(PRODUCT, ["SELECT count(*) FROM read_fastq('%s', gpus=1)" % path])
    assert rows(r[0]) == [["4000"]]
    _need(SQLRUN, PRODUCT)
    out = subprocess.run([SQLRUN], input=("LOAD '%s';\nSELECT count(*) FROM read_fastq('%s', gpus=%d);\n" % (PRODUCT, path, n_dev + 1)).encode(),
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    res = [json.loads(l) for l in out.stdout.decode().splitlines()]
    assert not res[1]["ok"] and "CUDA device" in res[1]["error"]


@pytest.mark.gpu
def test_one_pipeline_per_gpu_gives_the_same_answers(cuda_device, tmp_path):
    """read_fastq / read_fasta with one device pipeline per GPU (byte-range shards of the file): same rows in the same
    order as a single pipeline.  On a one-GPU box the shards all run on device 0 (gpus=1 pipelines are forced with
    EXON_B200_FORCE_READERS), which exercises the same shard logic."""
    import torch
    from oracle import oracle as O
    text, _ = util.random_fastq(37, 20000, min_len=30, max_len=150, tricky=False)
    path = _write(tmp_path, "mg.fastq", text)
    ftext, _ = util.random_fasta(39, 500, min_len=0, max_len=4000, tricky=False)
    fpath = _write(tmp_path, "mg.fasta", ftext)
    ref = O.parse_fastq(text)
    names = [x.decode() for x in ref.strings("name")]
    n_dev = torch.cuda.device_count()
    shards = 4
    env = {"EXON_B200_CHUNK_BYTES": str(300_000), "EXON_B200_FORCE_READERS": str(shards)}
    mq = "list_avg(quality_score_string_to_list(quality_scores)) > 60"
    stmts = [
        "SELECT count(*) FROM read_fastq('%s')" % path,
        "SELECT count(*), sum(length(sequence)), sum(length(quality_scores)) FROM read_fastq('%s') WHERE %s" % (path, mq),
        "CREATE TABLE t AS SELECT name, gc_content(sequence) AS g FROM read_fastq('%s')" % path,
        "SELECT name FROM t WHERE rowid IN (0, 1, 4999, 5000, 9999, 10000, 15000, 19999) ORDER BY rowid",
        "SELECT count(*), sum(g::DOUBLE) FROM t",
        "SELECT name FROM read_fastq('%s') LIMIT 4 OFFSET 12345" % path,
        "SELECT count(*), sum(length(sequence)), avg(gc_content(sequence)) FROM read_fasta('%s')" % fpath,
        "SELECT id FROM read_fasta('%s') LIMIT 3 OFFSET 250" % fpath,
    ]
    one = run_sql(PRODUCT, stmts, threads=4, env={"EXON_B200_CHUNK_BYTES": str(300_000)})
    many = run_sql(PRODUCT, stmts, threads=4, env=env)
    for a, b in zip(one, many):
        assert rows(a) == rows(b)
    assert rows(many[0]) == [["20000"]]
    assert [r[0] for r in rows(many[3])] == [names[i] for i in (0, 1, 4999, 5000, 9999, 10000, 15000, 19999)]
    assert [r[0] for r in rows(many[5])] == names[12345:12349]
    if n_dev >= 2:  # real devices
        real = run_sql(PRODUCT, [s.replace("')", "', gpus=%d)" % n_dev) for s in stmts], threads=n_dev, env={"EXON_B200_CHUNK_BYTES": str(300_000)})
        for a, b in zip(one, real):
            assert rows(a) == rows(b)


@pytest.mark.gpu
def test_copy_to_fastq_replays_the_reference_copy_test(tmp_path):
    # test/sql/exondb-release-with-deb-info/test_fastq_copy.test, statement by statement (the reference keeps it commented
    # out since it removed its writers); COPY returns the number of records written
    from oracle import oracle as O
    T = str(tmp_path)
    src = "%s/test.fastq" % G
    r = run_sql(PRODUCT, [
        "COPY (SELECT * FROM read_fastq('%s')) TO '%s/test.fastq' (FORMAT 'fastq')" % (src, T),
        "COPY (FROM read_fastq('%s')) TO '%s/test.fastq.gz' (FORMAT 'fastq')" % (src, T),
        "COPY (SELECT * FROM read_fastq('%s')) TO '%s/test.fastq.zst' (FORMAT 'fastq')" % (src, T),
        "COPY (SELECT * FROM read_fastq('%s')) TO '%s/test.fastq.gzip' (FORMAT 'fastq', COMPRESSION 'gzip')" % (src, T),
        "COPY (SELECT * FROM read_fastq('%s')) TO '%s/test.fastq.gzip' (FORMAT 'fastq', COMPRESSION 'gzip', FORCE true)" % (src, T),
        "SELECT COUNT(*) FROM read_fastq('%s/test.fastq.gzip', compression='gzip')" % T,
        "COPY (SELECT * FROM read_fastq('%s')) TO '%s/test.fastq.zstd' (FORMAT 'fastq', COMPRESSION 'zstd')" % (src, T),
        "SELECT COUNT(*) FROM read_fastq('%s/test.fastq.zstd', compression='zstd')" % T,
        "SELECT COUNT(*) FROM read_fastq('%s/test.fastq.gz')" % T,
        "SELECT COUNT(*) FROM '%s/test.fastq.zst'" % T,
        "SELECT * FROM read_fastq('%s/test.fastq') EXCEPT SELECT * FROM read_fastq('%s')" % (T, src),
    ])
    assert [int(scalar(x)) for x in r[:5]] == [2, 2, 2, 2, 2]
    assert int(scalar(r[5])) == 2 and int(scalar(r[6])) == 2 and int(scalar(r[7])) == 2 and int(scalar(r[8])) == 2 and int(scalar(r[9])) == 2
    assert rows(r[10]) == []
    # the reference's fixture is in canonical form, so the plain copy reproduces it byte for byte -- and equals the oracle's writer
    text = open(src, "rb").read()
    t = O.parse_fastq(text)
    want = O.format_fastq(t.strings("name"), t.strings("description"), t.strings("sequence"), t.strings("quality_scores"))
    assert open(T + "/test.fastq", "rb").read() == want == text


@pytest.mark.gpu
def test_copy_to_fasta_replays_the_reference_copy_test(tmp_path):
    # test_fasta_copy.test: plain, .gz, .zst, FORCE, explicit COMPRESSION, and the mixed NULL-description round trip (:74-86)
    from oracle import oracle as O
    T = str(tmp_path)
    src = "%s/test.fasta" % G
    r = run_sql(PRODUCT, [
        "COPY (SELECT * FROM read_fasta('%s')) TO '%s/test.fasta' (FORMAT 'fasta')" % (src, T),
        "SELECT COUNT(*) FROM read_fasta('%s/test.fasta')" % T,
        "COPY (SELECT * FROM read_fasta('%s')) TO '%s/test.fasta.gz' (FORMAT 'fasta')" % (src, T),
        "SELECT COUNT(*) FROM read_fasta('%s/test.fasta.gz')" % T,
        "COPY (SELECT * FROM read_fasta('%s')) TO '%s/test.fasta.zst' (FORMAT 'fasta')" % (src, T),
        "COPY (SELECT * FROM read_fasta('%s')) TO '%s/test.fasta.zst' (FORMAT 'fasta', FORCE true)" % (src, T),
        "COPY (SELECT * FROM read_fasta('%s')) TO '%s/test.fasta.zst' (FORMAT 'fasta')" % (src, T),
        "SELECT COUNT(*) FROM read_fasta('%s/test.fasta.zst')" % T,
        "COPY (SELECT * FROM read_fasta('%s')) TO '%s/test.fasta.gzip' (FORMAT 'fasta', COMPRESSION 'gzip')" % (src, T),
        "SELECT COUNT(*) FROM read_fasta('%s/test.fasta.gzip', compression='gzip')" % T,
        "COPY (FROM read_fasta('%s/test.mixed-desc.fasta')) TO '%s/test.mixed-desc.fasta' (FORMAT 'fasta')" % (G, T),
        "FROM read_fasta('%s/test.mixed-desc.fasta') WHERE description IS NULL" % T,
    ])
    assert int(scalar(r[0])) == 2 and int(scalar(r[1])) == 2 and int(scalar(r[2])) == 2 and int(scalar(r[3])) == 2 and int(scalar(r[4])) == 2 and int(scalar(r[5])) == 2
    assert not r[6]["ok"] and "exists" in r[6]["error"]  # "Now don't force it, and expect an error"
    assert int(scalar(r[7])) == 2 and int(scalar(r[8])) == 2 and int(scalar(r[9])) == 2 and int(scalar(r[10])) == 2
    assert rows(r[11]) == [["b", None, "ATCG"]]
    t = O.parse_fasta(open(src, "rb").read())
    assert open(T + "/test.fasta", "rb").read() == O.format_fasta(t.strings("id"), t.strings("description"), t.strings("sequence"))


@pytest.mark.gpu
def test_copy_filters_and_rewrites_a_larger_file(tmp_path):
    # the job the writers exist for: quality-filter a FASTQ into a new file.  The filter runs on the device inside the scan,
    # the passing rows come back as DuckDB vectors and are formatted on the device again; order is the file's order.
    from oracle import oracle as O
    text, recs = util.random_fastq(77, 6000, min_len=1, max_len=250, tricky=False)
    src = tmp_path / "in.fastq"
    src.write_bytes(text)
    for threads in (1, 4):
        out = tmp_path / ("out%d.fastq" % threads)
        r = run_sql(PRODUCT, ["COPY (SELECT * FROM read_fastq('%s') WHERE list_avg(quality_score_string_to_list(quality_scores)) > 45) "
                              "TO '%s' (FORMAT 'fastq')" % (src, out)], threads=threads)
        keep = [x for x in recs if O.mean_quality_pass(x[3], ">", 45.0)]
        assert int(scalar(r[0])) == len(keep) and 0 < len(keep) < 6000
        want = O.format_fastq([x[0] for x in keep], [x[1] for x in keep], [x[2] for x in keep], [x[3] for x in keep])
        assert out.read_bytes() == want


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [PRODUCT, REFGLUE], ids=["product", "refglue"])
def test_bgzf_input_through_sql(cuda_device, tmp_path, ext):
    """bgzip'ed FASTQ through DuckDB: the members are inflated on the device behind read_fastq, the
    compression parameter and the replacement scan -- same rows as the plain file, also through the reference's own glue."""
    from oracle import oracle as O
    bgzf_bytes = util.bgzf_bytes
    text, _ = util.random_fastq(17, 6000, min_len=10, max_len=150, tricky=False)
    path = _write(tmp_path, "b.fastq.gz", bgzf_bytes(text, sizes=[65280, 3, 40000]))
    other = _write(tmp_path, "b.dat", bgzf_bytes(text, block=30000))
    ref = O.parse_fastq(text)
    want = [tuple(None if v is None else v.decode() for v in row) for row in ref.rows()]
    mq = sum(O.mean_quality_pass(q, ">", 50.0) for q in ref.strings("quality_scores"))
    res = run_sql(ext, [
        "SELECT count(*) FROM read_fastq('%s')" % path,
        "SELECT * FROM read_fastq('%s')" % path,
        "SELECT count(*) FROM read_fastq('%s', compression='gzip')" % other,
        "SELECT count(*) FROM '%s' WHERE list_avg(quality_score_string_to_list(quality_scores)) > 50" % path,
    ], env={"EXON_B200_CHUNK_BYTES": str(150_000)})
    assert rows(res[0]) == [[str(ref.n)]]
    assert [tuple(r) for r in rows(res[1])] == want
    assert rows(res[2]) == [[str(ref.n)]]
    assert rows(res[3]) == [[str(mq)]]


@pytest.mark.gpu
def test_gpu_stats_rows_describe_the_scans(cuda_device, tmp_path):
    """exon_gpu_stats(): one row per reader that ran, newest first -- file, format, compression, rows, bytes."""
    bgzf_bytes = util.bgzf_bytes
    text, _ = util.random_fastq(19, 3000, min_len=10, max_len=150, tricky=False)
    fa, _ = util.random_fasta(21, 50, min_len=10, max_len=2000, tricky=False)
    pq = _write(tmp_path, "s.fastq.gz", bgzf_bytes(text))
    pa = _write(tmp_path, "s.fasta", fa)
    res = run_sql(PRODUCT, [
        "SELECT count(*) FROM read_fastq('%s')" % pq,
        "SELECT count(*), sum(length(sequence)) FROM read_fasta('%s')" % pa,
        "SELECT path, format, compression, io_path, failed, rows, file_bytes, bytes_done, blocks > 0, seconds_total > 0 FROM exon_gpu_stats()",
    ])
    got = rows(res[2])
    assert got[0] == [pa, "fasta", "none", "pinned blocks", "false", "50", str(len(fa)), str(len(fa)), "true", "true"]
    assert got[1][:6] == [pq, "fastq", "bgzf (inflated on the device)", "pinned blocks", "false", "3000"]
    assert got[1][6] == got[1][7] == str(os.path.getsize(pq))
