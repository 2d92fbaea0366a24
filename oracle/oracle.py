"""ctypes wrapper around oracle/liboracle.so (CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


class _Table(C.Structure):
    _fields_ = [
        ("n_records", C.c_int64),
        ("n_cols", C.c_int32),
        ("off", C.POINTER(C.c_int64) * 4),
        ("data", C.POINTER(C.c_uint8) * 4),
        ("cap", C.c_int64 * 4),
        ("desc_valid", C.POINTER(C.c_uint8)),
        ("rec_cap", C.c_int64),
        ("err_pos", C.c_int64),
        ("err", C.c_char * 128),
    ]


def build():
    src = os.path.join(_HERE, "exon_oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        u8p = C.c_char_p
        L.orc_parse_fastq.argtypes = [u8p, C.c_int64, C.POINTER(_Table)]
        L.orc_parse_fasta.argtypes = [u8p, C.c_int64, C.POINTER(_Table)]
        L.orc_free_table.argtypes = [C.POINTER(_Table)]
        L.orc_gc_content.argtypes = [u8p, C.c_int64]
        L.orc_gc_content.restype = C.c_float
        L.orc_gc_count.argtypes = [u8p, C.c_int64]
        L.orc_gc_count.restype = C.c_int64
        for f in (L.orc_reverse_complement, L.orc_complement, L.orc_transcribe, L.orc_reverse_transcribe, L.orc_translate):
            f.argtypes = [u8p, C.c_int64, C.c_char_p]
            f.restype = C.c_int64
        L.orc_quality_to_list.argtypes = [u8p, C.c_int64, C.POINTER(C.c_int32)]
        L.orc_quality_sum.argtypes = [u8p, C.c_int64]
        L.orc_quality_sum.restype = C.c_int64
        L.orc_mean_quality.argtypes = [u8p, C.c_int64]
        L.orc_mean_quality.restype = C.c_double
        L.orc_mean_quality_pass.argtypes = [u8p, C.c_int64, C.c_int, C.c_double]
        L.orc_mean_quality_pass.restype = C.c_int
        L.orc_fastq_count_mean_quality.argtypes = [
            C.c_void_p, C.c_int64, C.c_int, C.c_double,
            C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_fastq_count_mean_quality.restype = C.c_int64
        pp = C.POINTER(C.c_void_p)
        L.orc_format_fastq.argtypes = [pp, pp, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_format_fastq.restype = C.c_int64
        L.orc_format_fasta.argtypes = [pp, pp, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
        L.orc_format_fasta.restype = C.c_int64
        _lib = L
    return _lib


class OracleError(ValueError):
    def __init__(self, msg, pos):
        super().__init__("%s at byte %d" % (msg, pos))
        self.pos = pos
        self.msg = msg


class Table:
    """Parsed records as Arrow-style columns (numpy copies)."""

    def __init__(self, names, offsets, data, desc_valid):
        self.names = names
        self.offsets = offsets      # list of int64 arrays, n+1
        self.data = data            # list of uint8 arrays
        self.desc_valid = desc_valid
        self.n = len(desc_valid)

    def column(self, name):
        i = self.names.index(name)
        return self.offsets[i], self.data[i]

    def strings(self, name):
        off, dat = self.column(name)
        b = dat.tobytes()
        out = [b[off[i]:off[i + 1]] for i in range(self.n)]
        if name == "description":
            out = [s if v else None for s, v in zip(out, self.desc_valid)]
        return out

    def rows(self):
        cols = [self.strings(nm) for nm in self.names]
        return list(zip(*cols))


def _parse(fn, names, buf):
    buf = bytes(buf)
    t = _Table()
    rc = fn(buf, len(buf), C.byref(t))
    try:
        if rc != 0:
            raise OracleError(t.err.decode(), t.err_pos)
        n = t.n_records
        offs, dats = [], []
        for c in range(t.n_cols):
            off = np.ctypeslib.as_array(t.off[c], shape=(n + 1,)).copy()
            nb = int(off[n])
            dat = np.ctypeslib.as_array(t.data[c], shape=(nb,)).copy() if nb else np.zeros(0, np.uint8)
            offs.append(off)
            dats.append(dat)
        valid = np.ctypeslib.as_array(t.desc_valid, shape=(n,)).copy().astype(bool) if n else np.zeros(0, bool)
        return Table(names, offs, dats, valid)
    finally:
        lib().orc_free_table(C.byref(t))


def parse_fastq(buf):
    return _parse(lib().orc_parse_fastq, ["name", "description", "sequence", "quality_scores"], buf)


def parse_fasta(buf):
    return _parse(lib().orc_parse_fasta, ["id", "description", "sequence"], buf)


def gc_content(s):
    """float32 result as numpy.float32; None -> None (DuckDB NULL handling)."""
    if s is None:
        return None
    s = bytes(s)
    return np.float32(lib().orc_gc_content(s, len(s)))


def gc_count(s):
    s = bytes(s)
    return int(lib().orc_gc_count(s, len(s)))


class InvalidInput(ValueError):
    pass


def _lut(fn, s):
    if s is None:
        return None
    s = bytes(s)
    out = C.create_string_buffer(len(s) + 1)
    bad = fn(s, len(s), out)
    if bad >= 0:
        raise InvalidInput("Invalid character in sequence: %s" % chr(s[bad]))
    return out.raw[:len(s)]


def reverse_complement(s):
    return _lut(lib().orc_reverse_complement, s)


def complement(s):
    return _lut(lib().orc_complement, s)


def transcribe(s):
    return _lut(lib().orc_transcribe, s)


def reverse_transcribe(s):
    return _lut(lib().orc_reverse_transcribe, s)


def translate_dna_to_aa(s):
    if s is None:
        return None
    s = bytes(s)
    out = C.create_string_buffer(len(s) // 3 + 1)
    bad = lib().orc_translate(s, len(s), out)
    if bad == -2:
        raise InvalidInput("Invalid sequence length: %d" % len(s))
    if bad >= 0:
        raise InvalidInput("Invalid codon: %s" % s[bad:bad + 3].decode("latin-1"))
    return out.raw[:len(s) // 3]


def quality_score_string_to_list(s):
    s = bytes(s)
    out = np.zeros(len(s), np.int32)
    lib().orc_quality_to_list(s, len(s), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def quality_sum(s):
    s = bytes(s)
    return int(lib().orc_quality_sum(s, len(s)))


def mean_quality(s):
    s = bytes(s)
    return float(lib().orc_mean_quality(s, len(s)))


OPS = {">": 0, ">=": 1, "<": 2, "<=": 3, "=": 4, "!=": 5}


def mean_quality_pass(s, op, c):
    s = bytes(s)
    return bool(lib().orc_mean_quality_pass(s, len(s), OPS[op], float(c)))


def fastq_count_mean_quality(buf, op, c):
    """Whole-buffer reference query: returns (pass, n_records, gc_total, len_total).

    `buf` may be bytes or a numpy uint8 array (no copy for the latter)."""
    if isinstance(buf, np.ndarray):
        ptr, n = buf.ctypes.data, buf.size
    else:
        buf = bytes(buf)
        ptr, n = C.cast(C.c_char_p(buf), C.c_void_p).value, len(buf)
    nr, gc, tot = C.c_int64(), C.c_int64(), C.c_int64()
    p = lib().orc_fastq_count_mean_quality(ptr, n, OPS[op], float(c), C.byref(nr), C.byref(gc), C.byref(tot))
    if p < 0:
        raise OracleError("parse error", -1)
    return int(p), nr.value, gc.value, tot.value


def pack_strings(strings):
    """list of bytes / None -> (int64 offsets[n + 1], uint8 data, bool valid[n]); None counts as empty and not valid."""
    lens = np.fromiter((0 if x is None else len(x) for x in strings), np.int64, len(strings))
    off = np.zeros(len(strings) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    data = np.frombuffer(b"".join(x for x in strings if x is not None), np.uint8).copy()
    if data.size == 0:
        data = np.zeros(4, np.uint8)
    valid = np.fromiter((x is not None for x in strings), np.uint8, len(strings))
    return off, data, valid


def _format(fasta, columns, line_width=80):
    """columns: lists of bytes (description entries may be None) -> the file image as bytes."""
    n_cols = 3 if fasta else 4
    assert len(columns) == n_cols
    n = len(columns[0])
    packed = [pack_strings(c) for c in columns]
    offs = (C.c_void_p * 4)(*[p[0].ctypes.data for p in packed] + [None] * (4 - n_cols))
    dats = (C.c_void_p * 4)(*[p[1].ctypes.data for p in packed] + [None] * (4 - n_cols))
    valid = packed[1][2]
    cap = sum(int(p[0][-1]) for p in packed) + 8 * n + (int(packed[2][0][-1]) // max(line_width, 1) + n if fasta else 0) + 16
    out = np.zeros(cap, np.uint8)
    if fasta:
        m = lib().orc_format_fasta(offs, dats, valid.ctypes.data, n, line_width, out.ctypes.data, cap)
    else:
        m = lib().orc_format_fastq(offs, dats, valid.ctypes.data, n, out.ctypes.data, cap)
    assert 0 <= m <= cap, (m, cap)
    return out[:m].tobytes()


def format_fastq(names, descriptions, sequences, qualities):
    return _format(False, [names, descriptions, sequences, qualities])


def format_fasta(ids, descriptions, sequences, line_width=80):
    return _format(True, [ids, descriptions, sequences], line_width)
