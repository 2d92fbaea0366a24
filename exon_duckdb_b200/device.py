"""Device-resident pipeline: torch tensors own the HBM, the kernels come from
libexon_b200.so through the C ABI.  torch is plumbing here (allocation, streams,
events); nothing in this module computes on tensors with torch ops.

All functions take/return CUDA tensors; `buf` is a uint8 tensor holding the file
image (allocate with `alloc_input` or any 16-byte aligned CUDA tensor).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import F_LINES, F_QUAL, F_SEQ, NO_POS, ExonError, check, lib

UINT64_MAX = 0xFFFFFFFFFFFFFFFF


class FormatError(ValueError):
    """Malformed FASTA/FASTQ input (the reference aborts the query)."""

    def __init__(self, msg, pos=-1):
        super().__init__(msg)
        self.pos = pos


class InvalidInput(ValueError):
    """duckdb::InvalidInputException of reverse_complement / complement."""


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    # data_ptr() of a zero-length view is NULL; the storage address is what the C ABI wants
    return C.c_void_p(t.untyped_storage().data_ptr() + t.storage_offset() * t.element_size())


def _empty(n, dtype, device):
    return torch.empty(max(int(n), 1), dtype=dtype, device=device)[: int(n)]


def alloc_input(n, device="cuda"):
    """uint8 buffer with slack so the kernels' 16-byte staging never leaves the allocation."""
    return torch.empty(int(n) + 64, dtype=torch.uint8, device=device)[: int(n)]


def to_device(data, device="cuda"):
    """bytes / numpy uint8 -> CUDA uint8 tensor (test and tooling helper)."""
    import numpy as np

    arr = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    buf = alloc_input(arr.size, device)
    if arr.size:
        buf.copy_(torch.from_numpy(arr.copy()))
    return buf


def workspace(n_bytes, device):
    return torch.empty(lib().exb_scan_workspace_bytes(int(n_bytes)), dtype=torch.uint8, device=device)


def fetch_result(ws):
    res = _lib.ScanResult()
    check(lib().exb_scan_result_fetch(_ptr(ws), C.byref(res), _stream()))
    return res


class FastqScan:
    """Outputs of one exb_fastq_scan launch (device tensors; scalars after `.result`)."""

    def __init__(self):
        self.buf = None
        self.n = 0
        self.begin = 0
        self.flags = 0
        self.wide = False
        self.line_end = None
        self.seq_len = self.gc = self.qual_len = self.qsum = None
        self.ws = None
        self._res = None

    @property
    def result(self):
        if self._res is None:
            self._res = fetch_result(self.ws)
        return self._res

    def validate(self):
        r = self.result
        if r.err_pos != NO_POS:
            raise FormatError("malformed FASTQ record at byte %d" % r.err_pos, r.err_pos)
        if r.total_lines % 4 != 0:
            raise FormatError("truncated FASTQ record: %d lines" % r.total_lines)
        if r.overflow:
            raise ExonError(_lib.ERR_CAPACITY, "output capacity exceeded")
        return int(r.total_lines // 4)


def fastq_scan(buf, flags=F_SEQ | F_QUAL, rec_cap=None, begin=0, n=None, max_lines=UINT64_MAX, out=None):
    """Launch the single-pass FASTQ scan (asynchronous).  `out` reuses a previous FastqScan's buffers."""
    n = buf.numel() if n is None else int(n)
    dev = buf.device
    s = out if out is not None else FastqScan()
    span = n - begin
    if rec_cap is None:
        rec_cap = span // 32 + 4096
    if out is None:
        s.wide = n >= 0xFFFFFFFF
        if flags & F_LINES:
            s.line_end = _empty(4 * rec_cap, torch.int64 if s.wide else torch.int32, dev)
        if flags & F_SEQ:
            s.seq_len = _empty(rec_cap, torch.int32, dev)
            s.gc = _empty(rec_cap, torch.int32, dev)
        if flags & F_QUAL:
            s.qual_len = _empty(rec_cap, torch.int32, dev)
            s.qsum = _empty(rec_cap, torch.int32, dev)
        # the workspace carries one 8-byte record per line between the scan's two kernels
        s.ws = torch.empty(lib().exb_fastq_workspace_bytes(span + 16, 4 * rec_cap + 4), dtype=torch.uint8, device=dev)
        s.rec_cap = rec_cap
    s.buf, s.n, s.begin, s.flags, s._res = buf, n, begin, flags, None
    check(lib().exb_fastq_scan(_ptr(buf), begin, n, 1, None, max_lines, flags, _ptr(s.line_end),
                               s.line_end.numel() if s.line_end is not None else 0, 1 if s.wide else 0,
                               _ptr(s.seq_len), _ptr(s.gc), _ptr(s.qual_len), _ptr(s.qsum), s.rec_cap,
                               _ptr(s.ws), s.ws.numel(), _stream()))
    return s


def fastq_scan_sync(buf, flags=F_SEQ | F_QUAL, **kw):
    """Scan, growing the per-record capacity to the hard bound if the estimate was too small."""
    s = fastq_scan(buf, flags, **kw)
    if s.result.overflow:
        kw.pop("rec_cap", None)
        span = (kw.get("n") or buf.numel()) - kw.get("begin", 0)
        s = fastq_scan(buf, flags, rec_cap=span // 4 + 16, **kw)
    return s


def fastq_filter(scan, n_records, preds, want_pass=False, agg=None, device_count=False):
    """Per-record predicates (ANDed) + aggregates over the passing records (asynchronous).

    Returns (agg int64[8] tensor: count, sum seq_len, sum gc, sum qsum, sum qual_len; pass uint8 tensor or None)."""
    dev = scan.buf.device
    arr, k = _lib.predicates(preds)
    if agg is None:
        agg = torch.empty(8, dtype=torch.int64, device=dev)
    pas = _empty(n_records, torch.uint8, dev) if want_pass else None
    check(lib().exb_fastq_filter(_ptr(scan.seq_len), _ptr(scan.gc), _ptr(scan.qual_len), _ptr(scan.qsum), n_records,
                                 arr, k, _ptr(pas), _ptr(agg), _ptr(scan.ws) if device_count else None, _stream()))
    return agg, pas


class FastqCount:
    """Outputs of one fused exb_fastq_scan_filter launch: `agg` (int64[8] device tensor) and the scan's result block."""

    def __init__(self):
        self.agg = None
        self.ws = None
        self.buf = None
        self.n = 0
        self._res = None

    @property
    def result(self):
        if self._res is None:
            self._res = fetch_result(self.ws)
        return self._res

    def validate(self):
        r = self.result
        if r.err_pos != NO_POS:
            # the fused flavour knows that a line start contradicts its phase, not where: the plain scan locates it
            fastq_scan_sync(self.buf, 0, n=self.n).validate()
            raise FormatError("malformed FASTQ record at byte %d" % r.err_pos, r.err_pos)
        if r.total_lines % 4 != 0:
            raise FormatError("truncated FASTQ record: %d lines" % r.total_lines)
        return int(r.total_lines // 4)


def fastq_scan_filter(buf, preds, n=None, out=None):
    """COUNT-style read_fastq + quality-line predicates in ONE kernel (asynchronous): nothing per record is written.

    preds: [(field, op, value)] with field 'mean_quality' or 'qual_len'.  Returns a FastqCount (reuse it through `out`)."""
    n = buf.numel() if n is None else int(n)
    dev = buf.device
    c = out if out is not None else FastqCount()
    if out is None:
        c.agg = torch.zeros(8, dtype=torch.int64, device=dev)
        c.ws = workspace(n + 16, dev)
    c._res = None
    c.buf, c.n = buf, n
    arr, k = _lib.predicates(preds)
    check(lib().exb_fastq_scan_filter(_ptr(buf), 0, n, 1, None, arr, k, _ptr(c.agg), 0, _ptr(c.ws), c.ws.numel(), _stream()))
    return c


def fastq_scan_totals(buf, n=None, ws=None):
    """Fused TOTALS flavour (C5): returns (int64[8] device tensor: [1] sum len, [2] sum #GC, [3] sequence lines,
    [5] sum round(gc_content * 2^32)), the scan result block).  One pass, nothing per record is written."""
    n = buf.numel() if n is None else int(n)
    dev = buf.device
    if ws is None:
        ws = torch.empty(lib().exb_fastq_workspace_bytes(n + 16, n // 24 + 16384), dtype=torch.uint8, device=dev)
    agg = torch.zeros(8, dtype=torch.int64, device=dev)
    check(lib().exb_fastq_scan_totals(_ptr(buf), 0, n, 1, None, _ptr(agg), 0, _ptr(ws), ws.numel(), _stream()))
    return agg, fetch_result(ws)


def exclusive_scan_u32(x, n=None):
    n = x.numel() if n is None else n
    out = torch.empty(n + 1, dtype=torch.int64, device=x.device)
    ws = workspace(4 * n + 16, x.device)
    check(lib().exb_exclusive_scan_u32(_ptr(x), n, _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


def exclusive_scan_u32_multi(x, n, cols):
    """cols exclusive scans in one launch: x = uint32[cols * n] (column after column) -> int64[cols, n + 1]."""
    out = torch.empty((cols, n + 1), dtype=torch.int64, device=x.device)
    ws = workspace(4 * n * cols + 16, x.device)
    check(lib().exb_exclusive_scan_u32_multi(_ptr(x), n, cols, n, _ptr(out), n + 1, _ptr(ws), ws.numel(), _stream()))
    return out


def select_rows(pas):
    n = pas.numel()
    off = torch.empty(n + 1, dtype=torch.int64, device=pas.device)
    sel = _empty(n, torch.int64, pas.device)
    ws = workspace(4 * n + 16, pas.device)
    check(lib().exb_select_rows(_ptr(pas), n, _ptr(off), _ptr(sel), _ptr(ws), ws.numel(), _stream()))
    cnt = int(off[n].item())
    return sel[:cnt]


class Column:
    """Arrow-style string column on the device: int64 offsets (n+1), uint8 data, optional validity bytes."""

    def __init__(self, offsets, data, valid=None):
        self.offsets, self.data, self.valid = offsets, data, valid

    def __len__(self):
        return self.offsets.numel() - 1

    def to_pylist(self):
        off = self.offsets.cpu().numpy()
        dat = self.data.cpu().numpy().tobytes()
        val = self.valid.cpu().numpy() if self.valid is not None else None
        return [dat[off[i]:off[i + 1]] if (val is None or val[i]) else None for i in range(len(off) - 1)]


FASTQ_COLUMNS = ["name", "description", "sequence", "quality_scores"]
FASTA_COLUMNS = ["id", "description", "sequence"]


def fastq_table(buf, columns=None, preds=(), n=None, seq_map=None):
    """read_fastq on a device buffer: {column: Column}, projected to `columns`, filtered by `preds`.
    seq_map = 'reverse_complement' | 'complement': the sequence column comes out already mapped (fused into the gather,
    exb_fastq_gather_map); raises InvalidInput like the scalar function does."""
    columns = list(columns) if columns is not None else list(FASTQ_COLUMNS)
    dev = buf.device
    n = buf.numel() if n is None else int(n)
    flags = F_LINES | ((F_SEQ | F_QUAL) if preds else 0)
    scan = fastq_scan_sync(buf, flags, n=n)
    n_rec = scan.validate()
    sel = None
    n_rows = n_rec
    if preds:
        _, pas = fastq_filter(scan, n_rec, preds, want_pass=True)
        sel = select_rows(pas)
        n_rows = sel.numel()
    wide = 1 if scan.wide else 0
    if sel is None and n_rows > 0:
        return _fastq_split(buf, n, scan, n_rows, columns, seq_map)
    lens = _empty(4 * n_rows, torch.int32, dev)
    valid = _empty(n_rows, torch.uint8, dev)
    check(lib().exb_fastq_fields(_ptr(buf), 0, n, _ptr(scan.line_end), wide, _ptr(sel), n_rows, _ptr(lens), _ptr(valid), None, _ptr(scan.ws), _stream()))
    out = {}
    # Arrow offsets of all four columns in ONE launch, then ONE host round trip for the column sizes
    offs = exclusive_scan_u32_multi(lens, n_rows, 4)
    totals = offs[:, n_rows].cpu().tolist()
    for name in columns:
        c = FASTQ_COLUMNS.index(name)
        off, total = offs[c], totals[c]
        data = _empty(total, torch.uint8, dev)
        if name == "sequence" and seq_map is not None:
            mode = {"reverse_complement": _lib.MAP_REVERSE_COMPLEMENT, "complement": _lib.MAP_COMPLEMENT}[seq_map]
            bad = torch.empty(1, dtype=torch.int64, device=dev)
            check(lib().exb_fastq_gather_map(_ptr(buf), 0, n, _ptr(scan.line_end), wide, _ptr(sel), n_rows, c, _ptr(lens), _ptr(off),
                                             mode, _ptr(data), _ptr(bad), _stream()))
            b = int(bad.item())
            if b != -1:
                raise InvalidInput("Invalid character in sequence: %s" % chr(b & 0xFF))
        else:
            check(lib().exb_fastq_gather(_ptr(buf), 0, n, _ptr(scan.line_end), wide, _ptr(sel), n_rows, c, _ptr(lens), _ptr(off),
                                         _ptr(data), _stream()))
        out[name] = Column(off, data, valid if name == "description" else None)
    out["__n_rows__"] = n_rows
    return out


def _fastq_split(buf, n, scan, n_rows, columns, seq_map):
    """Unfiltered read_fastq: field split + offsets + all wanted columns in ONE launch (exb_fastq_split)."""
    dev = buf.device
    wide = 1 if scan.wide else 0
    mask = 0
    for name in columns:
        mask |= 1 << FASTQ_COLUMNS.index(name)
    offs = torch.empty((4, n_rows + 1), dtype=torch.int64, device=dev)
    valid = _empty(n_rows, torch.uint8, dev)
    scratch = torch.empty(lib().exb_fastq_split_scratch_bytes(n_rows), dtype=torch.uint8, device=dev)
    mode = -1
    bad = None
    if seq_map is not None:
        mode = {"reverse_complement": _lib.MAP_REVERSE_COMPLEMENT, "complement": _lib.MAP_COMPLEMENT}[seq_map]
        bad = torch.empty(1, dtype=torch.int64, device=dev)
    for attempt in range(2):
        # capacities: a guess that fits ordinary files, then the hard bound (no column is larger than the input)
        caps = [(n // 4 + 4096, n // 4 + 4096, n // 2 + 4096, n // 2 + 4096), (n + 16,) * 4][attempt]
        data = [torch.empty(caps[c] + 16, dtype=torch.uint8, device=dev) if (mask >> c) & 1 else None for c in range(4)]
        outs = (C.c_void_p * 4)(*[_ptr(d) for d in data])
        capv = (C.c_int64 * 4)(*[caps[c] if (mask >> c) & 1 else 0 for c in range(4)])
        check(lib().exb_fastq_split(_ptr(buf), 0, n, _ptr(scan.line_end), wide, n_rows, mask, _ptr(offs), _ptr(valid), outs, capv, _ptr(scratch),
                                    _ptr(scan.ws), mode, _ptr(bad), _stream()))
        if int(scratch[8:12].view(torch.int32).item()) == 0:
            break
    else:
        raise RuntimeError("exb_fastq_split: column capacity")
    if bad is not None:
        b = int(bad.item())
        if b != -1:
            raise InvalidInput("Invalid character in sequence: %s" % chr(b & 0xFF))
    totals = offs[:, n_rows].cpu().tolist()
    out = {}
    for name in columns:
        c = FASTQ_COLUMNS.index(name)
        out[name] = Column(offs[c], data[c][:totals[c]], valid if name == "description" else None)
    out["__n_rows__"] = n_rows
    return out


class FastaScan:
    def __init__(self):
        self.hdr_start = self.hdr_end = self.seq_off = self.gc_prefix = self.seq = None
        self.ws = None
        self._res = None

    @property
    def result(self):
        if self._res is None:
            self._res = fetch_result(self.ws)
        return self._res


def fasta_scan(buf, rec_cap=None, compact=True, n=None, seq_cap=None, out=None):
    """Launch the single-pass FASTA scan (+ sequence compaction) (asynchronous)."""
    n = buf.numel() if n is None else int(n)
    dev = buf.device
    s = out if out is not None else FastaScan()
    if out is None:
        if rec_cap is None:
            rec_cap = n // 64 + 4096
        s.rec_cap = rec_cap
        s.hdr_start = _empty(rec_cap, torch.int64, dev)
        s.hdr_end = _empty(rec_cap, torch.int64, dev)
        s.seq_off = _empty(rec_cap + 1, torch.int64, dev)
        s.gc_prefix = _empty(rec_cap + 1, torch.int64, dev)
        s.seq = alloc_input(n if seq_cap is None else seq_cap, dev) if compact else None
        s.ws = workspace(n + 16, dev)
    s.buf, s.n, s._res = buf, n, None
    check(lib().exb_fasta_scan(_ptr(buf), 0, n, 1, n, None, _ptr(s.hdr_start), _ptr(s.hdr_end), _ptr(s.seq_off), _ptr(s.gc_prefix),
                               s.rec_cap, _ptr(s.seq), s.seq.numel() if s.seq is not None else 0, _ptr(s.ws), s.ws.numel(),
                               _stream()))
    return s


def fasta_scan_sync(buf, **kw):
    s = fasta_scan(buf, **kw)
    if s.result.overflow:
        kw.pop("rec_cap", None)
        n = kw.get("n") or buf.numel()
        s = fasta_scan(buf, rec_cap=n // 2 + 16, **kw)
    r = s.result
    if r.err_pos != NO_POS:
        raise FormatError("malformed FASTA input at byte %d" % r.err_pos, r.err_pos)
    if r.overflow:
        raise ExonError(_lib.ERR_CAPACITY, "output capacity exceeded")
    return s


def fasta_table(buf, columns=None, n=None):
    """read_fasta on a device buffer: {column: Column} (+ '__gc__': per-record gc_content from the scan's prefixes)."""
    columns = list(columns) if columns is not None else list(FASTA_COLUMNS)
    dev = buf.device
    n = buf.numel() if n is None else int(n)
    s = fasta_scan_sync(buf, n=n, compact="sequence" in columns)
    n_rows = int(s.result.n_records)
    out = {"__n_rows__": n_rows, "__scan__": s}
    if "id" in columns or "description" in columns:
        lens = _empty(2 * n_rows, torch.int32, dev)
        starts = _empty(2 * n_rows, torch.int64, dev)
        valid = _empty(n_rows, torch.uint8, dev)
        err = torch.empty(1, dtype=torch.int64, device=dev)
        check(lib().exb_fasta_headers(_ptr(buf), n, _ptr(s.hdr_start), _ptr(s.hdr_end), n_rows, _ptr(lens), _ptr(starts), _ptr(valid),
                                      _ptr(err), _stream()))
        bad = int(err.item())
        if bad != -1:
            raise FormatError("FASTA definition without a name at byte %d" % bad, bad)
        offs = exclusive_scan_u32_multi(lens, n_rows, 2)
        totals = offs[:, n_rows].cpu().tolist()
        for c, name in enumerate(["id", "description"]):
            if name not in columns:
                continue
            ln = lens[c * n_rows:(c + 1) * n_rows] if n_rows else lens
            off, total = offs[c], totals[c]
            data = _empty(total, torch.uint8, dev)
            start = starts[c * n_rows:(c + 1) * n_rows] if n_rows else starts
            check(lib().exb_gather_ranges(_ptr(buf), _ptr(start), _ptr(ln), _ptr(off), n_rows, _ptr(data), total, _stream()))
            out[name] = Column(off, data, valid if c == 1 else None)
    if "sequence" in columns:
        out["sequence"] = Column(s.seq_off[:n_rows + 1], s.seq[:int(s.result.seq_bytes)])
    return out


def gc_from_prefix(seq_off, gc_prefix, n_rows):
    out = _empty(n_rows, torch.float32, seq_off.device)
    check(lib().exb_gc_from_prefix(_ptr(seq_off), _ptr(gc_prefix), n_rows, _ptr(out), _stream()))
    return out


def gc_from_counts(seq_len, gc, n_rows):
    out = _empty(n_rows, torch.float32, seq_len.device)
    check(lib().exb_gc_from_counts(_ptr(seq_len), _ptr(gc), n_rows, _ptr(out), _stream()))
    return out


def gc_content(col):
    """gc_content(VARCHAR) -> FLOAT over a device Column (sequence_functions/module.cpp:131-158, per-row formula)."""
    n = len(col)
    out = _empty(n, torch.float32, col.offsets.device)
    check(lib().exb_gc_content(_ptr(col.offsets), _ptr(col.data), n, _ptr(out), _stream()))
    return out


def _seq_map(col, mode):
    dev = col.offsets.device
    nb = col.data.numel()
    out = alloc_input(nb, dev)
    bad = torch.empty(1, dtype=torch.int64, device=dev)
    check(lib().exb_seq_map(_ptr(col.data), nb, mode, _ptr(out), _ptr(bad), _stream()))
    b = int(bad.item())
    if b != -1:
        ch = int(col.data[b].item())
        raise InvalidInput("Invalid character in sequence: %s" % chr(ch))
    return Column(col.offsets, out, col.valid)


def reverse_complement(col):
    """reverse_complement(VARCHAR) with the reference's semantics (module.cpp:30-69): A->C T->G C->A G->T, no reversal."""
    return _seq_map(col, _lib.MAP_REVERSE_COMPLEMENT)


def complement(col):
    return _seq_map(col, _lib.MAP_COMPLEMENT)


def transcribe(col):
    """transcribe(VARCHAR) (module.cpp:212-249): T -> U, A C G unchanged, anything else raises."""
    return _seq_map(col, _lib.MAP_TRANSCRIBE)


def reverse_transcribe(col):
    """reverse_transcribe(VARCHAR) (module.cpp:168-203): U -> T, A C G unchanged, anything else raises."""
    return _seq_map(col, _lib.MAP_REVERSE_TRANSCRIBE)


def translate_dna_to_aa(col):
    """translate_dna_to_aa(VARCHAR) (module.cpp:260-360) over a device Column; raises InvalidInput with the reference's
    message for the first offending row (length not a multiple of 3, or a codon outside the standard table)."""
    dev = col.offsets.device
    n = len(col)
    out = alloc_input(col.data.numel() // 3 + 1, dev)
    status = torch.empty(2, dtype=torch.int64, device=dev)
    check(lib().exb_translate(_ptr(col.offsets), _ptr(col.data), n, _ptr(out), _ptr(status), _stream()))
    bad_row, bad_pos = status.cpu().tolist()
    if bad_pos >= 0:
        raise InvalidInput("Invalid codon: %s" % bytes(col.data[bad_pos:bad_pos + 3].cpu().tolist()).decode("latin-1"))
    if bad_row >= 0:
        ln = int((col.offsets[bad_row + 1] - col.offsets[bad_row]).item())
        raise InvalidInput("Invalid sequence length: %d" % ln)
    base = col.offsets[0]
    off = torch.div(col.offsets - base, 3, rounding_mode="floor")
    return Column(off, out[:int(off[-1].item())] if n else out[:0], col.valid)


def quality_score_string_to_list(col):
    """LIST(INTEGER) child vector: int32 tensor of byte - 33; list offsets are col.offsets (fastq_functions/module.cpp:32-50)."""
    nb = col.data.numel()
    out = _empty(nb, torch.int32, col.offsets.device)
    check(lib().exb_quality_decode(_ptr(col.data), nb, _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------ writers (COPY ... TO (FORMAT 'fastq' | 'fasta'))
def _format(fasta, cols, line_width=80, out_cap=None):
    """Columns on the device -> (file image uint8 tensor, record offsets int64 tensor)."""
    n_cols = 3 if fasta else 4
    assert len(cols) == n_cols
    n = len(cols[0])
    dev = cols[0].offsets.device
    fc = _lib.FormatCols()
    for i, c in enumerate(cols):
        fc.d_off[i] = _ptr(c.offsets).value
        fc.d_data[i] = _ptr(c.data).value
    valid = cols[1].valid
    if valid is not None and valid.dtype != torch.uint8:
        valid = valid.to(torch.uint8)
    fc.d_desc_valid = _ptr(valid).value if valid is not None else None
    if out_cap is None:
        seq_bytes = int(cols[2].data.numel())
        out_cap = sum(int(c.data.numel()) for c in cols) + 8 * n + (seq_bytes // max(line_width, 1) + n if fasta else 0)
    out = alloc_input(out_cap, dev)
    row_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    scratch = torch.empty(lib().exb_format_scratch_bytes(n), dtype=torch.uint8, device=dev)
    if fasta:
        check(lib().exb_fasta_format(C.byref(fc), n, line_width, _ptr(row_off), _ptr(out), out_cap, _ptr(scratch), scratch.numel(), _stream()))
    else:
        check(lib().exb_fastq_format(C.byref(fc), n, _ptr(row_off), _ptr(out), out_cap, _ptr(scratch), scratch.numel(), _stream()))
    nb = C.c_int64(0)
    check(lib().exb_format_finish(_ptr(row_off), n, _ptr(scratch), out_cap, C.byref(nb), _stream()))
    return out[: nb.value], row_off


def fastq_format(name, description, sequence, quality_scores, out_cap=None):
    """The FASTQ file image of four device Columns (the inverse of fastq_table)."""
    return _format(False, [name, description, sequence, quality_scores], out_cap=out_cap)


def fasta_format(id, description, sequence, line_width=80, out_cap=None):
    """The FASTA file image of three device Columns, sequences wrapped at `line_width` (the inverse of fasta_table)."""
    return _format(True, [id, description, sequence], line_width, out_cap)


class Writer:
    """exb_writer_*: HOST columns -> file, formatted on the GPU (what the DuckDB copy function drives)."""

    def __init__(self, path, file_format, compression=None, force=False, device=0, line_width=None):
        self._w = C.c_void_p()
        check(lib().exb_writer_open(path.encode(), file_format.encode(), compression.encode() if compression else None, int(force), device,
                                    C.byref(self._w)))
        self.n_cols = 3 if file_format.lower() == "fasta" else 4
        if line_width is not None:
            check(lib().exb_writer_set_line_width(self._w, line_width))

    def append(self, columns, desc_valid=None):
        """columns: list of (int64 numpy offsets[n + 1], uint8 numpy data); desc_valid: uint8 numpy [n] or None."""
        import numpy as np

        assert len(columns) == self.n_cols
        n = len(columns[0][0]) - 1
        keep = [(np.ascontiguousarray(o, np.int64), np.ascontiguousarray(d, np.uint8)) for o, d in columns]
        offs = (C.c_void_p * 4)(*[o.ctypes.data for o, _ in keep] + [None] * (4 - self.n_cols))
        dats = (C.c_void_p * 4)(*[d.ctypes.data for _, d in keep] + [None] * (4 - self.n_cols))
        v = np.ascontiguousarray(desc_valid, np.uint8) if desc_valid is not None else None
        check(lib().exb_writer_append(self._w, n, offs, dats, v.ctypes.data if v is not None else None))

    def close(self):
        rows, nbytes = C.c_int64(0), C.c_int64(0)
        w, self._w = self._w, None
        check(lib().exb_writer_close(w, C.byref(rows), C.byref(nbytes)))
        return rows.value, nbytes.value


def bgzf_index(image, pos=0, max_out_bytes=1 << 62, max_blocks=1 << 20):
    """Member table of a BGZF file image (host bytes): exb_bgzf_index_host.  Returns (blocks, next_pos, out_bytes)."""
    import numpy as np

    arr = np.frombuffer(bytes(image), dtype=np.uint8)
    cap = min(max_blocks, arr.size // 26 + 2)
    tab = (_lib.BgzfBlock * cap)()
    n, nxt, outb = C.c_int64(), C.c_int64(), C.c_int64()
    check(lib().exb_bgzf_index_host(arr.ctypes.data, arr.size, pos, max_out_bytes, tab, cap, C.byref(n), C.byref(nxt), C.byref(outb)))
    return tab, n.value, nxt.value, outb.value


def bgzf_inflate(image, device="cuda", check_crc=True):
    """Inflate a whole BGZF file image on the device (exb_bgzf_index_host + exb_bgzf_inflate + exb_bgzf_finish);
    returns the text as a CUDA uint8 tensor.  Raises ExonError on a corrupt member."""
    import numpy as np

    tab, n, nxt, outb = bgzf_index(image)
    if nxt != len(image):
        raise _lib.ExonError(-3, "BGZF image has more members than the table holds")
    d_in = to_device(image, device)
    d_tab = torch.from_numpy(np.frombuffer(bytes(tab), dtype=np.uint8)[: max(n, 1) * 32].copy()).to(device)
    d_out = alloc_input(outb, device)
    d_state = torch.empty(16, dtype=torch.uint8, device=device)
    check(lib().exb_bgzf_inflate(_ptr(d_in), _ptr(d_tab), n, _ptr(d_out), _ptr(d_state), 1 if check_crc else 0, _stream()))
    bad = C.c_int64(-1)
    check(lib().exb_bgzf_finish(_ptr(d_state), C.byref(bad), _stream()))
    return d_out
