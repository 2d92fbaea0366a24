"""dist.py -- one FASTA/FASTQ file across ranks: contiguous byte ranges, a boundary-resync step at each shard edge,
and tiny collectives for what the shards must agree on.

One process per GPU.  Rank k owns bytes [lo_k, hi_k) = [k*N/G, (k+1)*N/G) of the file; a record belongs to the shard that
holds its first byte.  Data never moves between GPUs; what is exchanged is one 128-byte block per shard:

  FASTQ phase (exact)     '@' and '+' are legal quality characters, so a shard cannot tell from its own bytes which
                          of its lines are headers.  Every shard counts its newlines; the exclusive sum over the
                          earlier shards is the absolute index of its first line, and index mod 4 is the phase.
  aggregates              COUNT / sums are reduced with one all-reduce of 8 int64.

Two protocols are built on that exchange:

  ShardedFastqCount       COUNT-style queries, ONE pass over the bytes: the byte kernel (K1) of the fused scan does not
                          depend on the phase, so each shard runs it right away under a provisional predecessor; the
                          shards all-gather their result blocks, compose their true predecessor state (compose_prev)
                          and re-run only the light resolve kernel (exb_fastq_scan_filter_resolve).
  fastq_record_bounds /   Row-returning queries: the shards agree on record-aligned bounds S_0 <= ... <= S_G = N
  fasta_record_bounds     (S_k = first record start at or after lo_k); [S_k, S_{k+1}) is then a complete file image
                          for every other entry point of the library (tables, filters, projections).

The collectives go through a group object: TorchGroup (torch.distributed: NCCL on the GPU box, gloo in the CPU tests) or
LocalGroup, which holds G shards in one process.  The composition rules are pure functions over the gathered matrix,
so the CPU tests drive them with world_size 2 over gloo without a GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

# one row of shard state, int64 each (global file offsets)
LO, HI, NEWLINES, OPEN_START, TAIL_S, TAIL_G, OPEN_FLAGS, LS0 = 0, 1, 2, 3, 4, 5, 6, 7
N_LS = 5  # LS0..LS0+4: offsets of the first five line starts inside the shard (-1 = none)
STATE_WORDS = LS0 + N_LS
HALO = 128  # bytes of the file kept in front of a shard's first byte.  The scan only looks one byte back (CRLF); 128 keeps the
            # shard's first byte on the same 128-byte grid as its buffer, so that K1's TMA rows are whole 128-byte lines
            # (a shard whose tiles start 16 bytes off that grid scans ~5 % slower: profiles/round2_n2_launches.txt)
RESULT_WORDS = 16  # the 128-byte result block of a scan as int64 words
RECORD_WORDS = 32  # the 256-byte record of the single-exchange COUNT (exb_fastq_scan_filter_candidates)


def byte_range(n_bytes, rank, world):
    """[lo, hi) of rank's shard: contiguous, disjoint, covering [0, n_bytes)."""
    return n_bytes * rank // world, n_bytes * (rank + 1) // world


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers (first touch) and
    the threads that fill them are local to the GPU's PCIe root: with 8 ranks streaming 55 GB/s each, buffers on the
    wrong socket halve the end-to-end rate.  Returns the node, or None when the topology cannot be read."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            # the platform reports no affinity for the device (virtualised PCIe): with ONE node online every CPU and every
            # pinned page is local to every GPU -- there is nothing to bind, and that node is the answer
            with open("/sys/devices/system/node/online") as f:
                online = f.read().strip()
            if online.isdigit():
                return int(online)
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------ groups
class TorchGroup:
    """torch.distributed process group (NCCL: tensors on the rank's GPU; gloo: CPU tensors)."""

    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cpu")

    def all_gather_rows(self, row):
        """row: sequence of int64 or an int64 tensor on the group's device -> (world, len) numpy matrix, same on every rank."""
        t = row if hasattr(row, "numel") else self.torch.tensor([int(x) for x in row], dtype=self.torch.int64, device=self.device)
        t = t.contiguous().view(-1)
        out = self.torch.empty(self.world * t.numel(), dtype=self.torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view(self.world, t.numel()).cpu().numpy()

    def all_reduce_sum(self, t):
        """In-place sum (the tensor stays on its device: 64 bytes over NVLink for the aggregates)."""
        self.dist.all_reduce(t, group=self.group)
        return t


class PeerGroup(TorchGroup):
    """TorchGroup whose per-step exchange runs over NVLink peer memory with the library's own kernels
    (csrc/peer_exchange.cu) instead of NCCL: every rank owns a symmetric buffer (torch symmetric memory: cuMem
    allocations mapped into every peer of the box) and stores its 128-byte block / 64-byte aggregates straight into
    its peers' buffers.  torch.distributed still does the plumbing (rendezvous, the one-off exchange of the static
    shard ranges, barriers).  Raises if the ranks are not NVLink peers of one box -- callers fall back to TorchGroup."""

    def __init__(self, device, group=None):
        super().__init__(device, group)
        import torch.distributed._symmetric_memory as symm_mem

        torch, dist = self.torch, self.dist
        nbytes = int(_lib.lib().exb_peer_bytes())
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        assert self.handle.world_size == self.world and self.handle.rank == self.rank
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)  # nobody stores into a peer's buffer before every buffer has been zeroed
        self.d_peers = int(self.handle.buffer_ptrs_dev)
        self.seq = 0

    def blocks_view(self, seq):
        """The gathered result blocks of step `seq` in this rank's own buffer: int64[world * 16]."""
        off = int(_lib.lib().exb_peer_blocks_offset(seq))
        return self.buf[off:off + 128 * self.world].view(self.torch.int64)


class LocalGroup:
    """G shards held by one process (tests, multi-shard runs on one GPU): the "collective" completes once every shard
    has contributed its row."""

    def __init__(self, world):
        self.world = world
        self.rows = [None] * world

    def contribute(self, rank, row):
        self.rows[rank] = [int(x) for x in (row.tolist() if hasattr(row, "tolist") else row)]

    def matrix(self):
        assert all(r is not None for r in self.rows), "every shard must contribute before the exchange completes"
        return np.array(self.rows, dtype=np.int64)


# ------------------------------------------------------------------ pure composition rules
def state_row(lo, hi, begin, total_lines, open_line_start, tail_s, tail_g, pad, line_starts=()):
    """Shard state in GLOBAL file offsets from the scan's result block (local offsets: file byte lo = local `begin`)."""
    row = [0] * STATE_WORDS
    row[LO], row[HI], row[NEWLINES] = int(lo), int(hi), int(total_lines)
    row[OPEN_START] = int(lo) + (int(open_line_start) - int(begin))
    row[TAIL_S], row[TAIL_G], row[OPEN_FLAGS] = int(tail_s), int(tail_g), int(pad) & 3
    ls = [int(x) for x in line_starts][:N_LS]
    row[LS0:LS0 + N_LS] = ls + [-1] * (N_LS - len(ls))
    return row


def compose_prev(states, k, begin=HALO):
    """True predecessor state of shard k from the gathered shard states (pure).

    states[j] describes shard j scanned on its own: NEWLINES in [lo_j, hi_j); OPEN_START = offset of the first byte
    after its last newline (lo_j if it has none); TAIL_S / TAIL_G = byte sum / G,C count of the bytes after that;
    OPEN_FLAGS = '@'(2) / '+'(1) of the byte at OPEN_START when that byte lies in the shard.
    Returns an exb_scan_result in shard k's LOCAL coordinates (file byte lo_k sits at local offset `begin`), or None
    for shard 0, which has no predecessor."""
    if k == 0:
        return None
    lo_k = int(states[k][LO])
    tail_s = tail_g = 0
    start = int(states[0][LO])
    for j in range(k - 1, -1, -1):  # back to the last shard that saw a newline: the open line starts after it
        tail_s += int(states[j][TAIL_S])
        tail_g += int(states[j][TAIL_G])
        if int(states[j][NEWLINES]) > 0:
            start = int(states[j][OPEN_START])
            break
    flags = 0
    if start < lo_k:  # the shard that holds the open line's first byte knows what it is
        holder = next(j for j in range(k) if int(states[j][LO]) <= start < int(states[j][HI]))
        flags = int(states[holder][OPEN_FLAGS]) & 3
    # start == lo_k: the line starts with this shard's own first byte; the scan reads it itself (open_line_start == begin)
    res = _lib.ScanResult()
    res.total_lines = int(sum(int(states[j][NEWLINES]) for j in range(k)))
    res.open_line_start = begin + (start - lo_k)
    res.tail_s, res.tail_g, res.pad = tail_s, tail_g, flags
    res.err_pos = _lib.NO_POS
    return res


def fastq_record_bounds(states, n_bytes):
    """Record-aligned shard bounds S_0..S_G from the gathered states (pure).  S_k = offset of the first line that starts
    in [lo_k, hi_k) and whose absolute index is a multiple of 4; a shard without one is empty (S_k = S_{k+1})."""
    G = len(states)
    bounds = [0] * G + [int(n_bytes)]
    first = [None] * G
    lines_before = 0
    for k in range(G):
        lo, hi = int(states[k][LO]), int(states[k][HI])
        ls = [int(x) for x in states[k][LS0:LS0 + N_LS] if int(x) >= 0]
        if ls:
            # index of the first line that starts inside the shard: the line open at lo is `lines_before`; it starts
            # inside the shard only if it starts AT lo
            idx0 = lines_before if ls[0] == lo else lines_before + 1
            j = (-idx0) % 4
            if j < len(ls) and ls[j] < hi:
                first[k] = ls[j]
        lines_before += int(states[k][NEWLINES])
    for k in range(G - 1, -1, -1):
        bounds[k] = first[k] if first[k] is not None else bounds[k + 1]
    return bounds


def fasta_record_bounds(first_headers, n_bytes):
    """S_0..S_G for FASTA: first_headers[k] = offset of the first line-initial '>' in [lo_k, hi_k), or -1."""
    G = len(first_headers)
    bounds = [0] * G + [int(n_bytes)]
    for k in range(G - 1, -1, -1):
        bounds[k] = int(first_headers[k]) if int(first_headers[k]) >= 0 else bounds[k + 1]
    return bounds


# ------------------------------------------------------------------ device side
class Shard:
    """A shard's bytes in HBM: file bytes [lo - begin, hi) so that file byte lo sits at local offset `begin`
    (0 for the first shard, HALO for the others)."""

    def __init__(self, buf, lo, hi, begin, is_last):
        self.buf, self.lo, self.hi, self.begin, self.is_last = buf, int(lo), int(hi), int(begin), bool(is_last)
        self.n = self.begin + (self.hi - self.lo)

    @staticmethod
    def from_bytes(data, rank, world, device):
        """Cut rank's shard out of a host image of the whole file (tests / tools; a real host preads just this range)."""
        from . import device as D

        lo, hi = byte_range(len(data), rank, world)
        begin = 0 if rank == 0 else HALO
        halo = bytes(max(0, begin - lo)) + bytes(data[max(0, lo - begin):lo]) if begin else b""
        return Shard(D.to_device(halo + bytes(data[lo:hi]), device), lo, hi, begin, rank == world - 1)


def _result_block(ws):
    """The scan's 128-byte result block (first bytes of its workspace) viewed as 16 int64 words, on the device."""
    import torch

    return ws[:128].view(torch.int64)


def _state_from_block(words, lo, hi, begin, line_starts=()):
    # exb_scan_result: total_lines, open_line_start, err_pos, overflow|pad, n_records, seq_bytes, gc_total, tail_s, tail_g
    pad = (int(words[3]) >> 32) & 3
    return state_row(lo, hi, begin, int(words[0]), int(words[1]), int(words[7]), int(words[8]), pad, line_starts)


class ShardedFastqCount:
    """SELECT COUNT(*), sums FROM read_fastq(file) WHERE <quality-line predicates> over this rank's shard of the file.

    step() enqueues, with no host round trip: byte pass + line offsets + result block (exb_fastq_scan_filter_begin) ->
    exchange of the 128-byte result blocks -> exb_fastq_compose_prev (device) -> K2 under the true predecessor ->
    reduce of the aggregates.  `total` (int64[8],
    device) then holds the GLOBAL aggregates on every rank: [0] passing records, [3] sum of their Phred sums, [4] sum of
    their quality lengths, [6] lines of the file mod 4 (must be 0), [7] shards that met a malformed record (must be 0)."""

    def __init__(self, shard, preds, group, ranges=None):
        import torch

        from . import device as D

        self.shard, self.preds, self.group = shard, list(preds), group
        dev = shard.buf.device
        self.c = D.FastqCount()
        self.c.agg = torch.zeros(8, dtype=torch.int64, device=dev)
        self.c.ws = D.workspace(shard.n + 16, dev)
        self.c.buf, self.c.n = shard.buf, shard.n
        self.total = torch.zeros(8, dtype=torch.int64, device=dev)
        self.true_prev = torch.zeros(128, dtype=torch.uint8, device=dev)
        self.prov = None
        if shard.begin:  # provisional predecessor: "no line is open before `begin`"
            r = _lib.ScanResult()
            r.open_line_start = shard.begin
            r.err_pos = _lib.NO_POS
            self.prov = torch.zeros(128, dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib().exb_scan_result_store(D._ptr(self.prov), C.byref(r), D._stream()))
        self.arr, self.k = _lib.predicates(self.preds)
        # every shard's (lo, hi, begin): static, exchanged once
        self.ranges = np.asarray(ranges if ranges is not None else group.all_gather_rows([shard.lo, shard.hi, shard.begin]), dtype=np.int64)
        self.world = len(self.ranges)
        self.d_ranges = torch.from_numpy(self.ranges.copy()).to(dev)
        self.blocks = torch.zeros(self.world * RESULT_WORDS, dtype=torch.int64, device=dev)
        # single-exchange flavour: this shard's 256-byte record, and room for everybody's
        self.record = torch.zeros(RECORD_WORDS, dtype=torch.int64, device=dev)
        self.records = torch.zeros(self.world * RECORD_WORDS, dtype=torch.int64, device=dev)
        self.fused_exchange = os.environ.get("EXB_EXCHANGE_FUSED", "1") != "0"

    # -- the single-exchange flavour: K2 for all four phases BEFORE the exchange, then one all-gather of 256-byte records
    def scan_candidates(self):
        """K1 + line offsets + candidates kernel; returns the shard's record (int64[32], device)."""
        from . import device as D

        s = self.shard
        _lib.check(_lib.lib().exb_fastq_scan_filter_candidates(D._ptr(s.buf), s.begin, s.n, 1 if s.is_last else 0, D._ptr(self.prov), self.arr, self.k,
                                                               D._ptr(self.record), D._ptr(self.c.ws), self.c.ws.numel(), D._stream()))
        self.c._res = None
        return self.record

    def combine(self, records):
        """records: int64[world * 32] device tensor, every shard's record in shard order -> self.total (global aggregates)."""
        from . import device as D

        _lib.check(_lib.lib().exb_fastq_combine_records(D._ptr(records), D._ptr(self.d_ranges), self.world, self.arr, self.k, D._ptr(self.total),
                                                        D._stream()))
        return self.total

    # -- the phases (tests drive them shard by shard through a LocalGroup-style loop; step() chains them for a TorchGroup)
    def scan(self):
        """K1 + line offsets; returns the shard's result block (int64[16] view of the workspace, on the device)."""
        from . import device as D

        s = self.shard
        # K1 + line offsets + the result block; no K2 yet: the phase is only known after the exchange
        _lib.check(_lib.lib().exb_fastq_scan_filter_begin(D._ptr(s.buf), s.begin, s.n, 1 if s.is_last else 0, D._ptr(self.prov), self.arr, self.k,
                                                          D._ptr(self.c.ws), self.c.ws.numel(), D._stream()))
        self.c._res = None
        return _result_block(self.c.ws)

    def states_from(self, blocks):
        """Host view of gathered blocks as state rows (diagnostics / the host mirror compose_prev)."""
        return [_state_from_block(blocks[j], *self.ranges[j]) for j in range(len(blocks))]

    def resolve(self, blocks, rank):
        """blocks: int64[world * 16] device tensor, every shard's result block in shard order."""
        from . import device as D

        s = self.shard
        self.resolve_local(blocks, rank)
        blk = _result_block(self.c.ws)
        self.total.copy_(self.c.agg)
        self.total[7] = (blk[2] != 0).to(self.total.dtype)  # device err_pos word: 0 = no malformed record
        self.total[6] = (blk[0] & 3) if s.is_last else 0
        return self.total

    def resolve_local(self, blocks, rank):
        """compose + resolve kernels only (no aggregate bookkeeping): the peer-memory reduce does that on the device."""
        from . import device as D

        s = self.shard
        L = _lib.lib()
        prev = None
        if rank > 0:
            _lib.check(L.exb_fastq_compose_prev(D._ptr(blocks), D._ptr(self.d_ranges), self.world, rank, D._ptr(self.true_prev), D._stream()))
            prev = self.true_prev
        # K2 (bucket selection, 48 B per tile) under the true predecessor; the first shard has none
        _lib.check(L.exb_fastq_scan_filter_resolve(s.begin, s.n, 1 if s.is_last else 0, D._ptr(prev), self.arr, self.k,
                                                   D._ptr(self.c.agg), 0, D._ptr(self.c.ws), self.c.ws.numel(), D._stream()))
        self.c._res = None

    def step(self, after_scan=None, stage_marks=None):
        """One pass over the shard + the exchange; every launch is asynchronous on the current stream.
        after_scan: optional callable invoked between the scan and the exchange (bench.py records a CUDA event there);
        stage_marks: optional pair of callables invoked after the block exchange and after the resolve (stage timing)."""
        from . import device as D

        g = self.group
        if self.fused_exchange and stage_marks is None:
            # K1 + offsets + K2 for all four phases, then ONE exchange after which every rank holds the global aggregates
            rec = self.scan_candidates()
            if after_scan is not None:
                after_scan()
            if isinstance(g, PeerGroup):  # 6 launches of this library per step, one wait on NVLink flags, no NCCL
                g.seq += 1
                _lib.check(_lib.lib().exb_peer_count_fused(g.d_peers, g.rank, g.world, D._ptr(rec), D._ptr(self.d_ranges), self.arr, self.k, g.seq,
                                                           D._ptr(self.total), D._stream()))
                return self.total
            g.dist.all_gather_into_tensor(self.records, rec, group=g.group)
            return self.combine(self.records)
        blk = self.scan()
        if after_scan is not None:
            after_scan()
        if isinstance(g, PeerGroup):  # 9 launches of this library per step, no NCCL
            L = _lib.lib()
            g.seq += 1
            _lib.check(L.exb_peer_allgather_block(g.d_peers, g.rank, g.world, D._ptr(blk), g.seq, D._stream()))
            if stage_marks is not None:
                stage_marks[0]()
            self.resolve_local(g.blocks_view(g.seq), g.rank)
            if stage_marks is not None:
                stage_marks[1]()
            _lib.check(L.exb_peer_count_reduce(g.d_peers, g.rank, g.world, D._ptr(self.c.ws), D._ptr(self.c.agg), 1 if self.shard.is_last else 0,
                                               g.seq, D._ptr(self.total), D._stream()))
            return self.total
        g.dist.all_gather_into_tensor(self.blocks, blk, group=g.group)
        self.resolve(self.blocks, g.rank)
        g.all_reduce_sum(self.total)
        return self.total


def check_count(total):
    """Raise for a malformed / truncated file; returns the aggregates as a list of ints."""
    from .device import FormatError

    t = [int(x) for x in total.tolist()]
    if t[7] < 0:
        raise RuntimeError("peer-memory exchange timed out: a rank of the box did not answer")
    if t[7]:
        raise FormatError("malformed FASTQ record (in %d shard(s))" % t[7])
    if t[6]:
        raise FormatError("truncated FASTQ record: the file's line count is not a multiple of 4")
    return t


def fastq_shard_state(shard):
    """State row of a shard for fastq_record_bounds: newline count + the first N_LS line starts (one line-only scan, no
    per-record output; TAIL_S / TAIL_G are not computed by that flavour and must not be used from this row)."""
    import torch

    from . import device as D

    dev = shard.buf.device
    line_end = torch.empty(N_LS, dtype=torch.int64, device=dev)
    ws = torch.empty(_lib.lib().exb_fastq_workspace_bytes(shard.n + 16, shard.n // 24 + 16384), dtype=torch.uint8, device=dev)
    prov = None
    if shard.begin:
        r = _lib.ScanResult()
        r.open_line_start = shard.begin
        r.err_pos = _lib.NO_POS
        prov = torch.zeros(128, dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().exb_scan_result_store(D._ptr(prov), C.byref(r), D._stream()))
    for attempt in range(2):
        _lib.check(_lib.lib().exb_fastq_scan(D._ptr(shard.buf), shard.begin, shard.n, 0, D._ptr(prov), N_LS, _lib.F_LINES, D._ptr(line_end), N_LS, 1,
                                             None, None, None, None, 0, D._ptr(ws), ws.numel(), D._stream()))
        res = D.fetch_result(ws)
        if not res.overflow:
            break
        ws = torch.empty(_lib.lib().exb_fastq_workspace_bytes(shard.n + 16, shard.n + 1), dtype=torch.uint8, device=dev)
    k = min(N_LS, int(res.total_lines))
    ends = line_end[:k].cpu().tolist() if k else []
    prev_byte = int(shard.buf[shard.begin - 1].item()) if shard.begin and shard.lo > 0 else None
    ls = []
    if shard.hi > shard.lo and (shard.lo == 0 or prev_byte == 10):
        ls.append(shard.lo)
    for e in ends:
        g = shard.lo + (int(e) - shard.begin) + 1
        if g < shard.hi and len(ls) < N_LS:
            ls.append(g)
    return state_row(shard.lo, shard.hi, shard.begin, res.total_lines, res.open_line_start, res.tail_s, res.tail_g, res.pad, ls)


def fasta_first_header(shard, window=1 << 20):
    """Offset in the file of the first line-initial '>' in [lo, hi), or -1 (scans a growing window of the shard)."""
    import torch

    from . import device as D

    dev = shard.buf.device
    cap = 4096
    arrs = [torch.empty(cap + 1, dtype=torch.int64, device=dev) for _ in range(4)]
    b0 = shard.begin - 1 if shard.begin and shard.lo > 0 else shard.begin  # one byte back: is `lo` the start of a line?
    w = window
    while True:
        end = min(shard.n, shard.begin + w)
        ws = D.workspace(end + 16, dev)
        _lib.check(_lib.lib().exb_fasta_scan(D._ptr(shard.buf), b0, end, 1 if (end == shard.n and shard.is_last) else 0, shard.n, None,
                                             D._ptr(arrs[0]), D._ptr(arrs[1]), D._ptr(arrs[2]), D._ptr(arrs[3]), cap, None, 0,
                                             D._ptr(ws), ws.numel(), D._stream()))
        res = D.fetch_result(ws)
        k = min(cap, int(res.n_records))
        if k:
            hs = arrs[0][:k].cpu().numpy()
            hs = hs[hs >= shard.begin]
            if hs.size:
                return shard.lo + int(hs[0]) - shard.begin
        if end == shard.n:
            return -1
        w *= 4


class ShardedFastqTotals:
    """SELECT COUNT(*), SUM(#GC), SUM(length(sequence)), AVG(gc_content(sequence)) FROM read_fastq(file) over this rank's
    byte-range shard (BASELINE config C5).  Two-exchange protocol of ShardedFastqCount with the fused TOTALS flavour
    (fused=True, the default):

      byte pass with the sequence-line aggregates of all four phase hypotheses per tile + line offsets + result block
      (exb_fastq_scan_totals_begin) -> exchange of the 128-byte result blocks -> exb_fastq_compose_prev (device) ->
      K2 picks each tile's bucket under the true predecessor (exb_fastq_scan_totals_resolve) -> reduce.

    fused=False keeps the GENERAL scan flavour (exb_fastq_scan_begin -> ... -> exb_fastq_scan_resolve with
    EXB_F_SEQ | EXB_F_LOCAL_RECORDS -> exb_fastq_seq_totals over the per-record arrays): the two are bit-identical
    (tests/test_gpu_dist.py), the fused one writes nothing per record.

    A sequence line counts in the shard in which the LINE ends, so the sums are additive over shards whatever record the
    cut falls in.  `total` (int64[8], device) holds the GLOBAL values on every rank:
    [0] records (lines of the file / 4), [1] sum of sequence lengths, [2] sum of G/C, [5] sum of round(gc_content * 2^32),
    [6] lines of the file mod 4 (must be 0), [7] shards that met a malformed record (must be 0)."""

    def __init__(self, shard, group, ranges=None, rec_cap=None, max_lines=None, fused=None):
        import os

        import torch

        from . import device as D

        self.shard, self.group = shard, group
        self.fused = (os.environ.get("EXB_TOTALS_FUSED", "1") != "0") if fused is None else bool(fused)
        dev = shard.buf.device
        # capacities: estimates that fit any file with records of >= 32 bytes; a caller that knows its records passes tight ones
        self.ws = torch.empty(_lib.lib().exb_fastq_workspace_bytes(shard.n + 16, max_lines if max_lines else shard.n // 24 + 16384),
                              dtype=torch.uint8, device=dev)
        self.rec_cap = rec_cap if rec_cap else shard.n // 32 + 4096
        if not self.fused:
            self.seq_len = torch.zeros(self.rec_cap, dtype=torch.int32, device=dev)
            self.gc = torch.zeros(self.rec_cap, dtype=torch.int32, device=dev)
        self.agg = torch.zeros(8, dtype=torch.int64, device=dev)
        self.total = torch.zeros(8, dtype=torch.int64, device=dev)
        self.true_prev = torch.zeros(128, dtype=torch.uint8, device=dev)
        self.prov = None
        if shard.begin:  # provisional predecessor: "no line is open before `begin`"
            r = _lib.ScanResult()
            r.open_line_start = shard.begin
            r.err_pos = _lib.NO_POS
            self.prov = torch.zeros(128, dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib().exb_scan_result_store(D._ptr(self.prov), C.byref(r), D._stream()))
        self.ranges = np.asarray(ranges if ranges is not None else group.all_gather_rows([shard.lo, shard.hi, shard.begin]), dtype=np.int64)
        self.world = len(self.ranges)
        self.d_ranges = torch.from_numpy(self.ranges.copy()).to(dev)
        self.blocks = torch.zeros(self.world * RESULT_WORDS, dtype=torch.int64, device=dev)
        self.flags = _lib.F_SEQ | _lib.F_LOCAL_RECORDS

    def scan(self):
        from . import device as D

        s = self.shard
        if self.fused:
            _lib.check(_lib.lib().exb_fastq_scan_totals_begin(D._ptr(s.buf), s.begin, s.n, 1 if s.is_last else 0, D._ptr(self.prov),
                                                              D._ptr(self.ws), self.ws.numel(), D._stream()))
        else:
            _lib.check(_lib.lib().exb_fastq_scan_begin(D._ptr(s.buf), s.begin, s.n, 1 if s.is_last else 0, D._ptr(self.prov), self.flags,
                                                       D._ptr(self.ws), self.ws.numel(), D._stream()))
        return _result_block(self.ws)

    def resolve_local(self, blocks, rank):
        from . import device as D

        s = self.shard
        L = _lib.lib()
        prev = None
        if rank > 0:
            _lib.check(L.exb_fastq_compose_prev(D._ptr(blocks), D._ptr(self.d_ranges), self.world, rank, D._ptr(self.true_prev), D._stream()))
            prev = self.true_prev
        if self.fused:
            _lib.check(L.exb_fastq_scan_totals_resolve(s.begin, s.n, 1 if s.is_last else 0, D._ptr(prev), D._ptr(self.agg), 0,
                                                       D._ptr(self.ws), self.ws.numel(), D._stream()))
            return
        # a record that straddles the shard's first byte only gets the fields whose line ends here: start from zeros
        self.seq_len.zero_()
        self.gc.zero_()
        _lib.check(L.exb_fastq_scan_resolve(s.begin, s.n, 1 if s.is_last else 0, D._ptr(prev), D.UINT64_MAX, self.flags, None, 0, 0,
                                            D._ptr(self.seq_len), D._ptr(self.gc), None, None, self.rec_cap, D._ptr(self.ws), self.ws.numel(), D._stream()))
        self.agg.zero_()
        _lib.check(L.exb_fastq_seq_totals(D._ptr(self.seq_len), D._ptr(self.gc), self.rec_cap, D._ptr(self.agg), D._stream()))

    def finish_local(self):
        """agg[0] = lines this shard contributes to the record count (total lines of the file on the last shard), plus the
        bookkeeping words exb_peer_count_reduce / the NCCL path expect."""
        blk = _result_block(self.ws)
        if self.shard.is_last:
            self.agg[0] = blk[0] // 4  # total_lines of the last shard = lines of the whole file (its K2 ran under the true predecessor)

    def step(self):
        from . import device as D

        g = self.group
        blk = self.scan()
        if isinstance(g, PeerGroup):
            L = _lib.lib()
            g.seq += 1
            _lib.check(L.exb_peer_allgather_block(g.d_peers, g.rank, g.world, D._ptr(blk), g.seq, D._stream()))
            self.resolve_local(g.blocks_view(g.seq), g.rank)
            self.finish_local()
            _lib.check(L.exb_peer_count_reduce(g.d_peers, g.rank, g.world, D._ptr(self.ws), D._ptr(self.agg), 1 if self.shard.is_last else 0,
                                               g.seq, D._ptr(self.total), D._stream()))
            return self.total
        g.dist.all_gather_into_tensor(self.blocks, blk, group=g.group)
        self.resolve_local(self.blocks, g.rank)
        self.finish_local()
        blk = _result_block(self.ws)
        self.total.copy_(self.agg)
        self.total[7] = (blk[2] != 0).to(self.total.dtype)
        self.total[6] = (blk[0] & 3) if self.shard.is_last else 0
        g.all_reduce_sum(self.total)
        return self.total
