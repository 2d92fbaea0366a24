"""CPU-side checks of the C ABI: the library loads without a GPU, exports every
symbol include/exon_b200.h declares, and the host-side logic (file-type sniffing,
filter parsing, error reporting, generators, x87 emulation) behaves."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from tools import synth

from exon_duckdb_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cstr(p):
    return C.cast(p, C.c_char_p).value if p else None


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "exon_b200.h")).read()
    declared = set(re.findall(r"^EXB_API [^;(]*?\b(\w+)\s*\(", hdr, re.M))
    assert {"new_reader", "replacement_scan", "exb_fastq_scan", "exb_fasta_scan", "exb_gc_content", "exb_seq_map",
            "exb_quality_decode"} <= declared
    for name in declared:
        assert hasattr(L, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "python binding and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in L.exb_version()


@pytest.mark.parametrize("uri,want", [
    # test_fasta_scan.test:28-49 / test_fastq_scan.test:43-59: the suffixes the reference pins
    ("./x/test.fasta", b"FASTA"), ("./x/test.fasta.gz", b"FASTA"), ("./x/test.fastq", b"FASTQ"),
    ("./x/test.fastq.gz", b"FASTQ"), ("./x/test.fastq.zst", b"FASTQ"),
    ("a.fa", b"FASTA"), ("a.fna", b"FASTA"), ("a.fq.gz", b"FASTQ"),
    ("a.txt", None), ("a.gff", None), ("noext", None), ("a.gz", None),
])
def test_replacement_scan(uri, want):
    r = _lib.lib().replacement_scan(uri.encode())
    assert _cstr(r.file_type) == want


class _Stream(C.Structure):
    _fields_ = [("get_schema", C.c_void_p), ("get_next", C.c_void_p), ("get_last_error", C.c_void_p),
                ("release", C.c_void_p), ("private_data", C.c_void_p)]


def _new_reader(uri, fmt, filters=None, compression=None):
    s = _Stream()
    r = _lib.lib().new_reader(C.byref(s), uri.encode(), 2048, compression, fmt.encode(), filters)
    return s, _cstr(r.error)


def test_new_reader_errors_without_touching_the_gpu(golden_dir):
    # test_fasta_scan.test:51-53, test_fastq_scan.test:61-62: '' is an error at bind time
    for fmt in ("fasta", "fastq"):
        _, err = _new_reader("", fmt)
        assert err and b"could not register table" in err
    _, err = _new_reader(os.path.join(golden_dir, "test.fastq"), "gff")
    assert err and b"could not parse file_format" in err
    _, err = _new_reader(os.path.join(golden_dir, "test.fastq"), "fastq", b"nosuchcol='a'")
    assert err and b"unknown column" in err
    _, err = _new_reader(os.path.join(golden_dir, "test.fastq"), "fastq", b"name='a")
    assert err and b"unterminated" in err


def test_new_reader_schema_matches_reference_columns(golden_dir):
    pa = pytest.importorskip("pyarrow")
    for fmt, f, cols in (("fastq", "test.fastq", ["name", "description", "sequence", "quality_scores"]),
                         ("fasta", "test.fasta", ["id", "description", "sequence"])):
        s, err = _new_reader(os.path.join(golden_dir, f), fmt, b"description IS NOT NULL AND (%s='a' OR %s>='b')" % ((cols[0].encode(),) * 2))
        assert err is None
        rd = pa.RecordBatchReader._import_from_c(C.addressof(s))
        assert rd.schema.names == cols
        assert all(str(t) == "string" for t in rd.schema.types)  # VARCHAR, as GetArrowLogicalType maps "u"


def test_generators_are_deterministic_and_sized():
    L = _lib.lib()
    for kind, kw in (("illumina", {}), ("ont", dict(len_min=1000, len_max=3000)), ("fasta", dict(len_min=0, len_max=500)),
                     ("illumina", dict(crlf=True)), ("fasta", dict(len_min=10, len_max=200, crlf=True, wrap=7))):
        p = synth.gen_params(kind, 50, seed=7, **kw)
        n = synth.gen_size(p)
        a = np.zeros(n, np.uint8)
        b = np.zeros(n, np.uint8)
        assert synth.lib().exb_gen_host(C.byref(p), a.ctypes.data, n) == 0
        assert synth.lib().exb_gen_host(C.byref(p), b.ctypes.data, n) == 0
        assert (a == b).all() and a[-1] == 10
        # record i of a shard equals record first_record + i of the whole
        q = synth.gen_params(kind, 10, seed=7, first_record=40, **kw)
        m = synth.gen_size(q)
        c = np.zeros(m, np.uint8)
        assert synth.lib().exb_gen_host(C.byref(q), c.ctypes.data, m) == 0
        assert a[n - m:].tobytes() == c.tobytes()
        assert synth.lib().exb_gen_host(C.byref(p), a.ctypes.data, n - 1) == _lib.ERR_CAPACITY


def test_generated_inputs_parse_with_the_oracle():
    from oracle import oracle as O
    L = _lib.lib()
    p = synth.gen_params("illumina", 200, seed=20)
    n = synth.gen_size(p)
    a = np.zeros(n, np.uint8)
    synth.lib().exb_gen_host(C.byref(p), a.ctypes.data, n)
    t = O.parse_fastq(a.tobytes())
    assert t.n == 200
    assert all(len(s) == 150 for s in t.strings("sequence"))
    assert all(d in (b"1:N:0:ACGTACGT", b"2:N:0:ACGTACGT") for d in t.strings("description"))
    means = [O.mean_quality(q) for q in t.strings("quality_scores")]
    assert 20 < sum(means) / len(means) < 40 and any(m > 30 for m in means) and any(m <= 30 for m in means)
    p = synth.gen_params("fasta", 20, seed=3, len_min=100, len_max=5000)
    n = synth.gen_size(p)
    a = np.zeros(n, np.uint8)
    synth.lib().exb_gen_host(C.byref(p), a.ctypes.data, n)
    t = O.parse_fasta(a.tobytes())
    assert t.n == 20 and t.strings("id")[3] == b"contig3"


def test_x87_division_emulation_matches_long_double(tmp_path):
    """exb_x87_div (the device's stand-in for DuckDB's long double average) vs native x87 on this host."""
    src = tmp_path / "x87t.c"
    src.write_text(r'''
#include <stdio.h>
#include "%s/exon_duckdb_b200/csrc/x87div.h"
int main(void){ unsigned long long s=88172645463325252ULL; long bad=0,dr=0;
 for(long i=0;i<3000000;i++){ s^=s<<13; s^=s>>7; s^=s<<17; unsigned n=(unsigned)(s>>40)%%300+1; if(i%%3==0) n=(unsigned)(s>>33)|1;
  long long sum=(long long)((s>>8)%%(41ull*n+1)); if(i%%5==0) sum=-(long long)(s%%100000); if(i%%7==0) sum=(long long)(s>>12);
  double ref=(double)((long double)sum/(long double)n), got=exb_x87_div(sum,n); if(ref!=got) bad++; if(ref!=(double)sum/(double)n) dr++;
  if(exb_mean_cmp(sum,n,0,ref)!=0 || exb_mean_cmp(sum,n,1,ref)!=1) bad++; }
 printf("%%ld %%ld\n",bad,dr); return 0; }
''' % ROOT)
    exe = tmp_path / "x87t"
    subprocess.check_call(["gcc", "-O2", "-o", str(exe), str(src), "-lm"])
    bad, dr = subprocess.check_output([str(exe)]).split()
    assert int(bad) == 0
    assert int(dr) > 0  # the sample does contain double-rounding cases, so the emulation is exercised


# ---------------------------------------------------------------- host logic of the native reader (no GPU: nothing is scanned)
def _open2(path, fmt, filters=None, **kw):
    import ctypes as C
    from exon_duckdb_b200 import _lib
    h = C.c_void_p()
    o = _lib.reader_options(**kw)
    rc = _lib.lib().exb_reader_open2(str(path).encode(), fmt.encode(), None, 2048, filters, C.byref(o), C.byref(h))
    return rc, h, _lib.lib().exb_last_error()


def test_filter_text_is_parsed_at_open(golden_dir):
    """a4: the predicate text FilterToString emits is parsed when the reader is opened -- a malformed or unsupported predicate is
    the reference's 'could not execute sql' error, before any device work."""
    from exon_duckdb_b200 import _lib
    fq = os.path.join(golden_dir, "test.fastq")
    for good in (b"name = 'SEQ_ID'", b"(name = 'a' AND sequence != 'ACGT') OR description IS NULL", b"mean_quality(quality_scores)>30",
                 b"gc_content(sequence) <= 0.5 AND length(sequence) > 10"):
        rc, h, msg = _open2(fq, "fastq", good)
        assert rc == 0, (good, msg)
        _lib.lib().exb_reader_close(h)
    for bad, what in ((b"name = 'x' AND", b"expected a column or function name"), (b"nosuch(quality_scores) > 3", b"unsupported function"),
                      (b"(sequence = 'A'", b"expected ')'"), (b"quality_scores = 'II' extra", b"cannot parse filter")):
        rc, h, msg = _open2(fq, "fastq", bad)
        assert rc != 0 and b"could not execute sql" in msg and what in msg, (bad, msg)
    rc, h, msg = _open2(fq, "fastq", None, range_lo=10, range_hi=5)
    assert rc != 0 and b"bad byte range" in msg
    rc, h, msg = _open2(os.path.join(golden_dir, "test.fastq.gz"), "fastq", None, range_lo=0, range_hi=100)
    assert rc != 0 and b"one uncompressed file" in msg


def test_reader_plan_and_cache_state(golden_dir, tmp_path):
    """What a host needs to plan one reader per GPU (exb_reader_plan), and the cache-state / stats calls on a process that has not
    scanned anything."""
    import ctypes as C
    from exon_duckdb_b200 import _lib
    L = _lib.lib()

    def plan(uri, fmt, comp=None):
        total, n, sh = C.c_int64(), C.c_int32(), C.c_int32()
        rc = L.exb_reader_plan(str(uri).encode(), fmt.encode(), comp, C.byref(total), C.byref(n), C.byref(sh))
        return rc, total.value, n.value, sh.value

    fq = os.path.join(golden_dir, "test.fastq")
    assert plan(fq, "fastq") == (0, os.path.getsize(fq), 1, 1)
    rc, total, n, sh = plan(os.path.join(golden_dir, "test.fastq.gz"), "fastq")
    assert rc == 0 and n == 1 and sh == 0                     # compressed input is not cut into byte ranges
    rc, total, n, sh = plan(os.path.join(golden_dir, "fastq"), "fastq")
    assert rc == 0 and n == 2 and sh == 0                     # a directory shards by file
    assert plan(tmp_path / "missing.fastq", "fastq")[0] != 0
    assert L.exb_file_cache_state(fq.encode()) == 0           # never scanned
    assert L.exb_file_cache_state(str(tmp_path / "missing.fastq").encode()) < 0
    L.exb_stats_reset()
    st = (_lib.ScanStats * 4)()
    n = C.c_int32(-1)
    assert L.exb_stats_snapshot(st, 4, C.byref(n)) == 0 and n.value == 0


def test_ctypes_mirrors_of_the_round2_structs_match_the_header(tmp_path):
    """exb_scan_stats / exb_bgzf_block / exb_reader_options as a C compiler lays them out vs the ctypes mirrors in _lib.py."""
    import ctypes as C
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                   'int main(void) { printf("%%zu %%zu %%zu %%zu %%zu\\n", sizeof(exb_scan_stats), offsetof(exb_scan_stats, file_bytes), '
                   'offsetof(exb_scan_stats, seconds_total), sizeof(exb_bgzf_block), sizeof(exb_reader_options)); return 0; }\n'
                   % os.path.join(ROOT, "include", "exon_b200.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(_lib.ScanStats), _lib.ScanStats.file_bytes.offset, _lib.ScanStats.seconds_total.offset, C.sizeof(_lib.BgzfBlock),
                   C.sizeof(_lib.ReaderOptions)]
