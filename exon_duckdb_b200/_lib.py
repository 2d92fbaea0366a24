"""ctypes binding of libexon_b200.so (C ABI: include/exon_b200.h).

The CUDA library IS the product: if it is missing or cannot be loaded this
module raises -- there is no Python or CPU fallback for any operation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EXON_B200_LIB") or os.path.join(_HERE, "libexon_b200.so")  # override: debug builds only

# flags / enums (mirror include/exon_b200.h)
F_LINES, F_SEQ, F_QUAL, F_LOCAL_RECORDS = 1, 2, 4, 16
P_MEAN_QUALITY, P_GC_CONTENT, P_SEQ_LEN, P_QUAL_LEN = 0, 1, 2, 3
OPS = {">": 0, ">=": 1, "<": 2, "<=": 3, "=": 4, "==": 4, "!=": 5, "<>": 5}
MAP_REVERSE_COMPLEMENT, MAP_COMPLEMENT, MAP_TRANSCRIBE, MAP_REVERSE_TRANSCRIBE = 0, 1, 2, 3
ERR_CUDA, ERR_ARG, ERR_FORMAT, ERR_CAPACITY, ERR_IO, ERR_INVALID_CHAR = -1, -2, -3, -4, -5, -6
NO_POS = 0xFFFFFFFFFFFFFFFF


class ScanResult(C.Structure):
    _fields_ = [
        ("total_lines", C.c_uint64),
        ("open_line_start", C.c_int64),
        ("err_pos", C.c_uint64),
        ("overflow", C.c_uint32),
        ("pad", C.c_uint32),
        ("n_records", C.c_uint64),
        ("seq_bytes", C.c_uint64),
        ("gc_total", C.c_uint64),
        ("tail_s", C.c_int64),
        ("tail_g", C.c_int64),
        ("tail_hdr", C.c_uint64),
        ("crlf_lines", C.c_uint64),
    ]


class Predicate(C.Structure):
    _fields_ = [("field", C.c_int32), ("op", C.c_int32), ("value", C.c_double)]


class FormatCols(C.Structure):
    _fields_ = [("d_off", C.c_void_p * 4), ("d_data", C.c_void_p * 4), ("d_desc_valid", C.c_void_p)]


class BgzfBlock(C.Structure):
    _fields_ = [("in_off", C.c_int64), ("out_off", C.c_int64), ("clen", C.c_uint32), ("isize", C.c_uint32), ("crc32", C.c_uint32),
                ("reserved", C.c_uint32)]


class ScanStats(C.Structure):
    _fields_ = [("path", C.c_char * 256), ("n_files", C.c_int32), ("format", C.c_int32), ("device", C.c_int32), ("compression", C.c_int32),
                ("io_path", C.c_int32), ("failed", C.c_int32), ("file_bytes", C.c_int64), ("bytes_done", C.c_int64), ("rows", C.c_int64),
                ("blocks", C.c_int64), ("seconds_total", C.c_double), ("seconds_io", C.c_double), ("seconds_device", C.c_double),
                ("seconds_scan", C.c_double), ("seconds_select", C.c_double), ("seconds_materialise", C.c_double), ("seconds_first_block", C.c_double)]


class ReaderResult(C.Structure):
    _fields_ = [("error", C.c_void_p)]


class ReplacementScanResult(C.Structure):
    _fields_ = [("file_type", C.c_void_p)]


class ColumnView(C.Structure):
    _fields_ = [
        ("offsets", C.POINTER(C.c_int64)),
        ("data", C.POINTER(C.c_uint8)),
        ("valid", C.POINTER(C.c_uint8)),
        ("type", C.c_int32),
        ("kind", C.c_int32),
        ("strings", C.c_void_p),
        ("chunk_nulls", C.c_int64),
        ("valid_bits", C.POINTER(C.c_uint64)),
        ("values", C.c_void_p),
        ("list_entries", C.POINTER(C.c_uint64)),
        ("n_values", C.c_int64),
    ]


MAX_COMPUTED = 8
C_GC_CONTENT, C_SEQ_MAP, C_QUALITY_LIST, C_MEAN_QUALITY, C_SEQ_LENGTH, C_QUAL_LENGTH = 1, 2, 3, 4, 5, 6
RD_STRING_T, RD_NO_OFFSETS, RD_COPY_IO = 1, 2, 4
T_VARCHAR, T_FLOAT, T_DOUBLE, T_INT32_LIST, T_INT64 = 0, 1, 2, 3, 4


class Batch(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64),
        ("n_cols", C.c_int32),
        ("n_computed", C.c_int32),
        ("batch_index", C.c_int64),
        ("cols", ColumnView * (4 + MAX_COMPUTED)),
        ("owner", C.c_void_p),
    ]


class Computed(C.Structure):
    _fields_ = [("kind", C.c_int32), ("arg", C.c_int32)]


class ReaderOptions(C.Structure):
    _fields_ = [
        ("size", C.c_uint32),
        ("device", C.c_int32),
        ("range_lo", C.c_int64),
        ("range_hi", C.c_int64),
        ("file_lo", C.c_int32),
        ("file_hi", C.c_int32),
        ("column_mask", C.c_uint32),
        ("flags", C.c_uint32),
        ("n_computed", C.c_int32),
        ("computed", Computed * MAX_COMPUTED),
    ]


def reader_options(column_mask=0xF, device=-1, range_lo=0, range_hi=0, file_lo=0, file_hi=0, flags=0, computed=()):
    o = ReaderOptions()
    o.size = C.sizeof(ReaderOptions)
    o.device = device
    o.range_lo, o.range_hi = range_lo, range_hi
    o.file_lo, o.file_hi = file_lo, file_hi
    o.column_mask = column_mask
    o.flags = flags
    o.n_computed = len(computed)
    for i, (kind, arg) in enumerate(computed):
        o.computed[i].kind = kind
        o.computed[i].arg = arg
    return o


class ExonError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("exon_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


# every symbol include/exon_b200.h declares: (restype, argtypes)
_vp, _i64, _u64, _i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
SIGNATURES = {
    "exb_last_error": (C.c_char_p, []),
    "exb_version": (C.c_char_p, []),
    "exb_device_available": (_i32, []),
    "new_reader": (ReaderResult, [_vp, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p]),
    "replacement_scan": (ReplacementScanResult, [C.c_char_p]),
    "exb_free_string": (None, [_vp]),
    "exb_reader_open": (_i32, [C.c_char_p, C.c_char_p, C.c_char_p, _i64, C.c_char_p, C.c_uint32, C.POINTER(_vp)]),
    "exb_reader_open2": (_i32, [C.c_char_p, C.c_char_p, C.c_char_p, _i64, C.c_char_p, C.POINTER(ReaderOptions), C.POINTER(_vp)]),
    "exb_reader_progress": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "exb_reader_plan": (_i32, [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "exb_device_count": (_i32, []),
    "exb_reader_columns": (_i32, [_vp, C.POINTER(C.c_char_p), _i32]),
    "exb_reader_next": (_i32, [_vp, C.POINTER(Batch)]),
    "exb_batch_release": (None, [C.POINTER(Batch)]),
    "exb_reader_count": (_i32, [_vp, C.POINTER(_i64)]),
    "exb_stats_snapshot": (_i32, [C.POINTER(ScanStats), _i32, C.POINTER(_i32)]),
    "exb_stats_reset": (None, []),
    "exb_file_cache_state": (_i32, [C.c_char_p]),
    "exb_reader_io_path": (_i32, [_vp]),
    "exb_reader_close": (None, [_vp]),
    "exb_scan_workspace_bytes": (_i64, [_i64]),
    "exb_fastq_workspace_bytes": (_i64, [_i64, _i64]),
    "exb_fastq_scan": (_i32, [_vp, _i64, _i64, _i32, _vp, _u64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "exb_fastq_scan_filter": (_i32, [_vp, _i64, _i64, _i32, _vp, C.POINTER(Predicate), _i32, _vp, _i32, _vp, _i64, _vp]),
    "exb_fastq_scan_totals": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _i32, _vp, _i64, _vp]),
    "exb_fastq_scan_totals_begin": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _vp]),
    "exb_fastq_scan_totals_resolve": (_i32, [_i64, _i64, _i32, _vp, _vp, _i32, _vp, _i64, _vp]),
    "exb_fastq_scan_filter_begin": (_i32, [_vp, _i64, _i64, _i32, _vp, C.POINTER(Predicate), _i32, _vp, _i64, _vp]),
    "exb_fastq_scan_begin": (_i32, [_vp, _i64, _i64, _i32, _vp, _i32, _vp, _i64, _vp]),
    "exb_fastq_seq_totals": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "exb_fastq_scan_resolve": (_i32, [_i64, _i64, _i32, _vp, _u64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "exb_fastq_scan_filter_resolve": (_i32, [_i64, _i64, _i32, _vp, C.POINTER(Predicate), _i32, _vp, _i32, _vp, _i64, _vp]),
    "exb_fastq_scan_filter_candidates": (_i32, [_vp, _i64, _i64, _i32, _vp, C.POINTER(Predicate), _i32, _vp, _vp, _i64, _vp]),
    "exb_fastq_combine_records": (_i32, [_vp, _vp, _i32, C.POINTER(Predicate), _i32, _vp, _vp]),
    "exb_peer_count_fused": (_i32, [_vp, _i32, _i32, _vp, _vp, C.POINTER(Predicate), _i32, _u64, _vp, _vp]),
    "exb_fastq_compose_prev": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "exb_peer_bytes": (_i64, []),
    "exb_peer_blocks_offset": (_i64, [_u64]),
    "exb_peer_allgather_block": (_i32, [_vp, _i32, _i32, _vp, _u64, _vp]),
    "exb_peer_count_reduce": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _u64, _vp, _vp]),
    "exb_scan_result_store": (_i32, [_vp, C.POINTER(ScanResult), _vp]),
    "exb_scan_result_fetch": (_i32, [_vp, C.POINTER(ScanResult), _vp]),
    "exb_fastq_filter": (_i32, [_vp, _vp, _vp, _vp, _i64, C.POINTER(Predicate), _i32, _vp, _vp, _vp, _vp]),
    "exb_fastq_fields": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "exb_exclusive_scan_u32": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "exb_exclusive_scan_u32_multi": (_i32, [_vp, _i64, _i32, _i64, _vp, _i64, _vp, _i64, _vp]),
    "exb_select_rows": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    "exb_fastq_gather": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "exb_fastq_split_scratch_bytes": (_i64, [_i64]),
    "exb_fastq_split": (_i32, [_vp, _i64, _i64, _vp, _i32, _i64, C.c_uint32, _vp, _vp, C.POINTER(_vp), C.POINTER(_i64), _vp, _vp, _i32, _vp, _vp]),
    "exb_fastq_gather_map": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "exb_fasta_scan": (_i32, [_vp, _i64, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "exb_fasta_headers": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "exb_gather_ranges": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "exb_gc_from_prefix": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "exb_gc_from_counts": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "exb_gc_content": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "exb_seq_map": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "exb_translate": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "exb_quality_decode": (_i32, [_vp, _i64, _vp, _vp]),
    "exb_gc_content_host": (_i32, [_vp, _vp, _i64, _vp]),
    "exb_seq_map_host": (_i32, [_vp, _i64, _i32, _vp, C.POINTER(_i64)]),
    "exb_translate_host": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "exb_quality_decode_host": (_i32, [_vp, _i64, _vp]),
    "exb_format_scratch_bytes": (_i64, [_i64]),
    "exb_fastq_format": (_i32, [C.POINTER(FormatCols), _i64, _vp, _vp, _i64, _vp, _i64, _vp]),
    "exb_fasta_format": (_i32, [C.POINTER(FormatCols), _i64, _i32, _vp, _vp, _i64, _vp, _i64, _vp]),
    "exb_format_finish": (_i32, [_vp, _i64, _vp, _i64, C.POINTER(_i64), _vp]),
    "exb_writer_open": (_i32, [C.c_char_p, C.c_char_p, C.c_char_p, _i32, _i32, C.POINTER(_vp)]),
    "exb_writer_set_line_width": (_i32, [_vp, _i32]),
    "exb_writer_append": (_i32, [_vp, _i64, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "exb_writer_close": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "exb_fastq_count_host": (_i32, [_vp, _i64, C.POINTER(Predicate), _i32, _i64, _i32, C.POINTER(_i64), C.POINTER(ScanResult)]),
    "exb_engine_create": (_i32, [_i32, _i64, C.POINTER(_vp)]),
    "exb_engine_destroy": (None, [_vp]),
    "exb_engine_fastq_count": (_i32, [_vp, _vp, _i64, C.POINTER(Predicate), _i32, C.POINTER(_i64), C.POINTER(ScanResult)]),
    "exb_host_alloc": (_vp, [_i64]),
    "exb_host_free": (None, [_vp]),
    "exb_bgzf_probe_host": (_i32, [_vp, _i64]),
    "exb_bgzf_index_host": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(BgzfBlock), _i64, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "exb_bgzf_compress_bound": (_i64, [_i64]),
    "exb_bgzf_compress_host": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _i64, C.POINTER(_i64)]),
    "exb_bgzf_scratch_bytes": (_i64, []),
    "exb_bgzf_inflate": (_i32, [_vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "exb_bgzf_finish": (_i32, [_vp, C.POINTER(_i64), _vp]),
    "exb_bgzf_status": (_i32, [C.POINTER(C.c_uint32), C.POINTER(_i64)]),
}

_lib = None


def lib():
    """The loaded library; raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "exon_duckdb_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = ABI drift between header and library
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise ExonError(rc, lib().exb_last_error().decode("utf-8", "replace"))


def predicates(preds):
    """[(field, op, value)] with field one of 'mean_quality', 'gc_content', 'seq_len', 'qual_len' -> ctypes array."""
    names = {"mean_quality": P_MEAN_QUALITY, "gc_content": P_GC_CONTENT, "seq_len": P_SEQ_LEN, "qual_len": P_QUAL_LEN}
    arr = (Predicate * max(1, len(preds)))()
    for i, (f, op, v) in enumerate(preds):
        arr[i].field = names[f] if isinstance(f, str) else int(f)
        arr[i].op = OPS[op] if isinstance(op, str) else int(op)
        arr[i].value = float(v)
    return arr, len(preds)
