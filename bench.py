#!/usr/bin/env python
"""bench.py -- read_fastq + mean-quality filter throughput on B200 (BASELINE.json configs[1]).

Workload (per GPU): synthetic Illumina FASTQ, 20 M reads x 150 bp (~7 GB), query
    SELECT COUNT(*) FROM read_fastq(f) WHERE list_avg(quality_score_string_to_list(quality_scores)) > 30
One step = one pass of the hot path over the whole file image:
    exb_fastq_scan_filter: TMA-fed tile kernel (every byte once: newline masks, Phred sums, predicate per line into
    4 phase buckets) -> chain-free offset scan of the per-tile line counts (3 short launches) -> combine kernel (picks
    each tile's bucket)
`value`  : input already resident in HBM, CUDA events on the launching stream, max over ranks.
`e2e`    : the same query through the plugin's own C ABI (exb_reader_open2 + exb_reader_count + exb_reader_close, the
           calls the DuckDB extension's init_global makes for this statement) on a FILE (tmpfs), every step: the file's
           page cache (registered with CUDA by the reader after the file's first scan: pinned in place) -> H2D -> scan /
           filter kernels -> the count read back.  `e2e_first_scan` is the same call on the path a first scan takes
           (page cache -> pinned blocks -> H2D, EXB_RD_COPY_IO); `h2d_peak_gbs` (pinned cudaMemcpy, measured here, all
           ranks at once) is the bound both are compared with.  `e2e_pinned_image` keeps round 1's number: the
           host-buffer engine (exb_engine_fastq_count) fed from an already pinned image.
`paths`  : (N = 1) the other configurations of BASELINE.json, device-resident, each with its algorithmic GB/s, fraction
           of the HBM peak and committed DRAM traffic: C3 (wrapped FASTA, gc_content per contig), C4 (ONT reads,
           reverse_complement projection), C2 general scan and 4-column materialisation (tools/paths.py).
`--impl reference` : the reference's CPU implementation of the path.  Its scan cannot be built here
           (Rust crates absent), so this arm times the oracle's record-at-a-time port
           (oracle/exon_oracle.c) on all host cores, one shard of whole records per thread.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "read_fastq + mean-quality filter (COUNT) throughput, input file bytes"
UNIT = "GB/s"
READS = 20_000_000
READ_LEN = 150
SEED = 20
THRESH = 30.0


def workload_config(n_gpus, reads):
    return {
        "workload": "C2: synthetic Illumina FASTQ %d reads x %d bp per GPU, COUNT(*) WHERE mean quality > %g" % (reads, READ_LEN, THRESH),
        "reads_per_gpu": reads,
        "query": "SELECT COUNT(*) FROM read_fastq(f) WHERE list_avg(quality_score_string_to_list(quality_scores)) > 30",
        "sharding": "one file of n_gpus x reads_per_gpu records cut into byte ranges whose edges fall inside records; per step: scan under a "
                    "provisional phase, exchange of the 128-byte result blocks, device-side composition + resolve kernel "
                    "(exon_duckdb_b200/dist.py), reduce of the 8 int64 aggregates (over NVLink peer memory, else NCCL: see `exchange`); no data-path collective"
        if n_gpus > 1 else "single GPU",
        "l2": "input (~7 GB) is larger than L2 (126 MB); no flush needed between steps",
    }


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def traffic_per_launch(tr, n_bytes):
    """DRAM bytes of one launch of the dominant kernel at THIS input size.  The capture may have been taken on a
    smaller file of the same generator (ncu replays each kernel ~40x); the kernel streams, so traffic scales with the input."""
    if not tr or not tr.get("dram_bytes_per_launch"):
        return None
    cap = tr.get("input_bytes") or n_bytes
    return tr["dram_bytes_per_launch"] * (n_bytes / cap)


def run_reference(args):
    """CPU arm: the oracle's record-at-a-time port, one shard of whole records per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from exon_duckdb_b200 import _lib
    from oracle import oracle as O
    from tools import synth
    L = _lib.lib()
    O.lib()
    cores = max(1, min(os.cpu_count() or 1, 64))
    # bounded sample: ~1 s of work per thread per step
    per_thread = int(os.environ.get("EXB_REF_READS_PER_THREAD", "1500000"))
    shards = []
    for t in range(cores):
        p = synth.gen_params("illumina", per_thread, seed=SEED, first_record=t * per_thread, len_min=READ_LEN, len_max=READ_LEN)
        n = synth.gen_size(p)
        a = np.empty(n, np.uint8)
        shards.append((p, a, n))
    # generation is outside the timed region; do it threaded as well
    def gen(i):
        p, a, n = shards[i]
        synth.lib().exb_gen_host(C.byref(p), a.ctypes.data, n)
    ths = [threading.Thread(target=gen, args=(i,)) for i in range(cores)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    total_bytes = sum(s[2] for s in shards)
    results = [None] * cores

    def work(i):
        results[i] = O.fastq_count_mean_quality(shards[i][1], ">", THRESH)

    def step():
        ths = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        [t.start() for t in ths]
        [t.join() for t in ths]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = total_bytes / dt / 1e9
    reads = per_thread * cores
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args.gpus, READS),
        "reads_per_s": reads / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d reads (%d per thread x %d threads, %.2f GB) of the C2 generator per step; "
                                   "oracle/exon_oracle.c orc_fastq_count_mean_quality; the reference's own scan needs Rust crates that are absent" %
                                   (reads, per_thread, cores, total_bytes / 1e9)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pass": int(sum(r[0] for r in results)), "records": int(sum(r[1] for r in results)),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("EXB_BENCH_READS", READS)))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-paths", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--tmp", default=os.environ.get("EXB_BENCH_TMP", "/dev/shm"))
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from exon_duckdb_b200 import _lib, device as D
    from tools import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    from exon_duckdb_b200.dist import bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local)  # before any pinned allocation: host buffers land next to the GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # NCCL writes its version banner to fd 1 at communicator creation; stdout carries exactly one JSON line, so fd 1
        # points at stderr until that line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    # ---- synthetic input, generated on the device
    preds = [("mean_quality", ">", THRESH)]
    sharded = None
    if world == 1:
        p = synth.gen_params("illumina", args.reads, seed=SEED, first_record=0, len_min=READ_LEN, len_max=READ_LEN)
        buf = synth.gen_device(p, dev)
        n_bytes = buf.numel()
        # COUNT(*) + a predicate on the quality line: scan and filter run as ONE kernel (exb_fastq_scan_filter);
        # projection push-down means nothing per record is written at all.
        cnt = D.fastq_scan_filter(buf, preds)
        n_rec = cnt.validate()
        assert n_rec == args.reads
        agg = cnt.agg
        want_pass = None
        e2e_src = buf

        def step(timers=None):
            if timers is not None:
                timers[0].record()
            D.fastq_scan_filter(buf, preds, out=cnt)
            if timers is not None:
                timers[1].record()
    else:
        # ONE file of world x reads records, cut into byte ranges whose edges fall INSIDE records (about 100 bytes past a
        # record start), so every step runs the boundary-resync protocol of exon_duckdb_b200/dist.py: scan under a
        # provisional phase, all-gather of the 128-byte result blocks, device-side composition, resolve kernel, all-reduce.
        from exon_duckdb_b200 import dist as XD

        exchange = "nvlink peer memory (symmetric buffers), ONE exchange per step: exb_fastq_scan_filter_candidates + exb_peer_count_fused"
        try:
            if os.environ.get("EXB_EXCHANGE", "peer") != "peer":
                raise RuntimeError("EXB_EXCHANGE=%s" % os.environ["EXB_EXCHANGE"])
            grp = XD.PeerGroup(dev)
            ok = 1
        except Exception as ex:  # not NVLink peers of one box / symmetric memory unavailable
            sys.stderr.write("rank %d: peer-memory exchange unavailable (%s); using NCCL\n" % (rank, ex))
            grp, ok = None, 0
        okt = torch.tensor([ok], dtype=torch.int64, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0:  # all ranks take the same path
            grp = XD.TorchGroup(dev)
            exchange = "nccl all-gather (256 B per shard) + local combine (exb_fastq_combine_records)"

        def build_shard(R):
            """This rank's byte range of ONE file of world x R records whose cuts fall INSIDE records."""
            def rec_size(i):
                q = synth.gen_params("illumina", 1, seed=SEED, first_record=i, len_min=READ_LEN, len_max=READ_LEN)
                return synth.gen_size(q)

            def delta(k):  # offset of shard k's first byte inside record k*R (96 .. 223 bytes into a ~350-byte record)
                # the shard's device buffer (HALO bytes before that byte) starts on a 128-byte boundary of the generated
                # image, as a per-GPU allocation of a host application would: 16 bytes is what the C ABI requires, but
                # K1's TMA rows are 128 bytes and a buffer that straddles them costs ~5 % (profiles/round2_n2_launches.txt)
                if k == 0:
                    return 0
                s_prev = rec_size(k * R - 1)
                return 96 + ((XD.HALO - s_prev - 96) % 128)

            first = rank * R - (1 if rank else 0)
            count = R + (1 if rank else 0) + (1 if rank < world - 1 else 0)
            p = synth.gen_params("illumina", count, seed=SEED, first_record=first, len_min=READ_LEN, len_max=READ_LEN)
            gbuf = synth.gen_device(p, dev)
            s_prev = rec_size(rank * R - 1) if rank else 0
            s_next = rec_size((rank + 1) * R) if rank < world - 1 else 0
            own = gbuf.numel() - s_prev - s_next  # bytes of records [rank*R, (rank+1)*R)
            owns = grp.all_gather_rows([own])[:, 0]
            off0 = int(owns[:rank].sum())  # file offset of record rank*R
            lo = off0 + delta(rank)
            hi = off0 + own + (delta(rank + 1) if rank < world - 1 else 0)
            begin = XD.HALO if rank else 0
            local0 = s_prev + delta(rank) - begin  # view: file byte lo sits at local offset `begin`
            assert local0 % 128 == 0 and local0 >= 0
            buf = gbuf[local0:local0 + begin + (hi - lo)]
            return XD.Shard(buf, lo, hi, begin, rank == world - 1), gbuf, s_prev, own

        R = args.reads
        shard, gbuf, s_prev, own = build_shard(R)
        buf = shard.buf
        lo, hi = shard.lo, shard.hi
        n_bytes = hi - lo
        sharded = XD.ShardedFastqCount(shard, preds, grp)
        # cross-check: the same records counted shard-locally on whole-record ranges
        whole = gbuf[s_prev:s_prev + own]
        if (s_prev % 16) != 0:
            whole = D.to_device(whole.cpu().numpy(), dev)
        chk = D.fastq_scan_filter(whole, preds)
        assert chk.validate() == R
        want = chk.agg.clone()
        dist.all_reduce(want)
        want_pass = int(want[0].item())
        del chk
        e2e_src = gbuf[s_prev:s_prev + own]  # the rank's whole records: the file image a host application would hold
        agg = sharded.total

        def step(timers=None):
            if timers is not None:
                timers[0].record()
            # scan -> exchange of the result blocks -> compose + resolve -> reduce of COUNT / sums (dist.py)
            sharded.step(after_scan=timers[1].record if timers is not None else None,
                         stage_marks=(timers[2].record, timers[3].record) if timers is not None and len(timers) > 2 else None)
            if timers is not None and len(timers) > 2:
                timers[4].record()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    n_ev = 5 if (world > 1 and os.environ.get("EXB_BENCH_STAGES")) else 2  # stage timing: scan | exchange | resolve | reduce
    scan_ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(n_ev)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        # the barrier is the LAST thing before the timed region: NVML start-up and event creation take a different time
        # on every rank, and in a lock-step exchange a rank that starts 2 ms late is charged to every other rank's total
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(scan_ev[i])
    e1.record()
    torch.cuda.synchronize()
    sampler.stop()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    scan_ms = sum(ev[0].elapsed_time(ev[1]) for ev in scan_ev) / args.steps
    if n_ev == 5:
        st = [sum(ev[i].elapsed_time(ev[i + 1]) for ev in scan_ev) / args.steps for i in range(4)]
        gap = (ms_total - sum(ev[0].elapsed_time(ev[4]) for ev in scan_ev)) / args.steps
        sys.stderr.write("rank %d stages (ms/step): scan %.4f  block exchange %.4f  compose+resolve %.4f  reduce %.4f  between steps %.4f\n"
                         % (rank, st[0], st[1], st[2], st[3], gap))
    got = agg.cpu().tolist()
    tmax = torch.tensor([ms_total, scan_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, scan_ms = tmax.tolist()
    ms_step = ms_total / args.steps
    tb = torch.tensor([n_bytes], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tb)
    total_bytes = int(tb.item())
    value = total_bytes / (ms_step * 1e-3) / 1e9
    n_pass = got[0]  # after the all-reduce this is already the global count
    if sharded is not None:
        XD.check_count(agg)
        assert n_pass == want_pass, (n_pass, want_pass)

    # ---- end to end: file (tmpfs) -> the plugin's reader C ABI -> the count, every step
    e2e = None
    e2e_first = None
    e2e_pinned = None
    host_ptr = None
    if not args.no_e2e:
        e2e_bytes = e2e_src.numel()
        host_ptr = L.exb_host_alloc(e2e_bytes)
        if not host_ptr:
            raise SystemExit("exb_host_alloc failed")
        host = np.ctypeslib.as_array(C.cast(host_ptr, C.POINTER(C.c_uint8)), shape=(e2e_bytes,))
        torch.from_numpy(host).copy_(e2e_src)  # untimed: put the file image where a host application would have it
        tmp_dir = args.tmp if os.path.isdir(args.tmp) and os.access(args.tmp, os.W_OK) else "/tmp"
        path = os.path.join(tmp_dir, "exb_bench_rank%d.fastq" % rank)
        host.tofile(path)  # untimed: the input FILE of this rank (its whole records)
        # pinned H2D peak of this box, all ranks copying at once: the bound of any end-to-end number
        pin = torch.from_numpy(host[:min(e2e_bytes, 1 << 30)])
        dst = torch.empty(pin.numel(), dtype=torch.uint8, device=dev)
        if world > 1:
            dist.barrier()
        h2d_peak = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(pin, non_blocking=True)
            torch.cuda.synchronize()
            h2d_peak = max(h2d_peak, pin.numel() / (time.perf_counter() - t0) / 1e9)
        del dst
        tp = torch.tensor([h2d_peak], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp)  # sum over ranks = what the box delivered with every GPU copying
        h2d_peak_all = tp.item()

        e2e_steps = max(3, min(args.steps, 10))
        # round 1's number, kept for comparison: the host-buffer engine fed from an already pinned image
        eng = C.c_void_p()
        _lib.check(L.exb_engine_create(local, 64 << 20, C.byref(eng)))
        parr, k = _lib.predicates(preds)
        eagg = (C.c_int64 * 8)()
        for _ in range(2):
            _lib.check(L.exb_engine_fastq_count(eng, host_ptr, e2e_bytes, parr, k, eagg, None))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _lib.check(L.exb_engine_fastq_count(eng, host_ptr, e2e_bytes, parr, k, eagg, None))
        dt = (time.perf_counter() - t0) / e2e_steps
        assert eagg[5] == args.reads, eagg[5]
        n_pinned = int(eagg[0])
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
        e2e_pinned = {"value": total_bytes / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "api": "exb_engine_fastq_count(pinned host image)"}
        L.exb_engine_destroy(eng)
        if world > 1:  # the pinned image is not needed any more (N ranks x 7 GB of it next to N x 7 GB of tmpfs files)
            L.exb_host_free(host_ptr)
            host_ptr = None
            del host, pin

        filt = ("mean_quality(quality_scores)>%r" % THRESH).encode()
        opt = _lib.reader_options(column_mask=0, device=local)

        def reader_count(flags=0):
            opt.flags = flags
            h = C.c_void_p()
            _lib.check(L.exb_reader_open2(path.encode(), b"fastq", None, 2048, filt, C.byref(opt), C.byref(h)))
            n = C.c_int64()
            rc = L.exb_reader_count(h, C.byref(n))
            direct = L.exb_reader_io_path(h)
            L.exb_reader_close(h)
            _lib.check(rc)
            return n.value, direct

        def timed_reader(flags):
            for _ in range(2):
                reader_count(flags)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                n, direct = reader_count(flags)
            dt = (time.perf_counter() - t0) / e2e_steps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return n, direct, tt.item()

        n_chunks = (e2e_bytes + (64 << 20) - 1) // (64 << 20)
        api = "exb_reader_open2(file, filters='mean_quality(quality_scores)>30') + exb_reader_count + exb_reader_close"

        def e2e_entry(n, dt, note):
            return {"value": total_bytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes,
                    "d2h_bytes_per_step": n_chunks * (C.sizeof(_lib.ScanResult) + 16 + 8) + 8,
                    "ms_per_step": dt * 1e3, "steps": e2e_steps, "pass": int(n),
                    "h2d_peak_gbs": h2d_peak_all, "frac_of_h2d_peak": (total_bytes / dt / 1e9) / h2d_peak_all if h2d_peak_all else None,
                    "api": api, "file": path, "note": note}

        # (1) the path a FIRST scan of a file takes: page cache -> pinned blocks (16 host threads) -> PCIe -> scan
        n_e2e, _, dt = timed_reader(_lib.RD_COPY_IO)
        e2e_first = e2e_entry(n_e2e, dt, "host wall clock around open + count + close, max over ranks; the file lives on tmpfs (page cache), every "
                                         "byte is copied into pinned blocks, sent over PCIe and scanned inside the timed region "
                                         "(EXB_RD_COPY_IO: what the first scan of a file does)")
        # (2) every later scan: the first complete scan had the file's page cache registered with CUDA (pinned in place), so
        # the bytes go page cache -> PCIe -> scan with no host copy.  Nothing is cached on the device: all of the file
        # crosses PCIe and is scanned inside the timed region of every step.
        reader_count(0)
        t_w = time.perf_counter()
        while L.exb_file_cache_state(path.encode()) == 1 and time.perf_counter() - t_w < 120:
            time.sleep(0.01)
        t_reg = time.perf_counter() - t_w
        n2, direct, dt2 = timed_reader(0)
        assert n2 == n_e2e == n_pinned, (n2, n_e2e, n_pinned)
        if direct:
            e2e = e2e_entry(n2, dt2, "host wall clock around open + count + close, max over ranks; repeated scan of a tmpfs file whose page "
                                     "cache the reader registered with CUDA after its first scan (cudaHostRegister, %.2f s in the background): "
                                     "every byte goes page cache -> PCIe -> scan inside the timed region, no host copy; "
                                     "`e2e_first_scan` is the same call on the copy path" % t_reg)
            e2e["io_path"] = "registered page cache (DMA from the file's pages)"
            e2e_first["io_path"] = "page cache -> pinned blocks -> DMA"
        else:  # this kernel / file system refuses to pin page-cache pages: the copy path is the only one
            e2e = e2e_first
            e2e["io_path"] = "page cache -> pinned blocks -> DMA (registration not available here)"
            e2e_first = None
        try:
            os.unlink(path)
        except OSError:
            pass

    # ---- CPU baseline beside it: oracle port, 1 thread, bounded sample of the same bytes (rank 0 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        sample_reads = min(args.reads, 8_000_000)
        if host_ptr:
            # cut the sample at a record boundary: the records are generated in order, so re-measure its size
            q = synth.gen_params("illumina", sample_reads, seed=SEED, first_record=0, len_min=READ_LEN, len_max=READ_LEN)
            sample_bytes = synth.gen_size(q) if sample_reads < args.reads else n_bytes
            sample = host[:sample_bytes]
        else:
            sample = buf[:0].cpu().numpy()
            sample_bytes = 0
        if sample_bytes:
            t0 = time.perf_counter()
            r = O.fastq_count_mean_quality(sample, ">", THRESH)
            dt = time.perf_counter() - t0
            cpu = {"value": sample_bytes / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "first %d reads (%.2f GB) of the same workload, oracle/exon_oracle.c record-at-a-time port, 1 thread, %.1f s"
                             % (sample_reads, sample_bytes / 1e9, dt),
                   "pass": int(r[0])}
    if host_ptr:
        L.exb_host_free(host_ptr)

    # ---- the other BASELINE configurations, device-resident (N = 1)
    path_rows = None
    if world == 1 and not args.no_paths:
        from tools import paths as P
        rep = P.Report(verbose=False)
        P.c2_paths(rep, buf, args.reads, iters=5, full=False)
        buf = cnt = None
        torch.cuda.empty_cache()
        P.c3_paths(rep, dev, int(os.environ.get("EXB_BENCH_CONTIGS", "6000")), 500_000, iters=5)   # C3: 3 Gbp wrapped at 60
        P.c4_paths(rep, dev, int(os.environ.get("EXB_BENCH_ONT_READS", "200000")), iters=3)       # C4: 200 k ONT reads, ~12 GB
        P.bgzf_paths(rep, dev, iters=3)                                                           # bgzip'ed FASTQ inflated on the device
        path_rows = rep.rows

    # ---- C5 (BASELINE configs[4]): ~100 GB of the same FASTQ as ONE file cut into byte ranges over the N GPUs (cuts inside
    # records); SELECT COUNT(*), SUM(#GC), SUM(length(sequence)), AVG(gc_content(sequence)); device-resident, strong scaling
    c5 = None
    if not args.no_c5:
        from exon_duckdb_b200 import dist as XD
        # release everything C2 held on the device (closures above see the same cells)
        buf = cnt = e2e_src = None
        if world > 1:
            gbuf = sharded = shard = whole = None
        torch.cuda.empty_cache()
        total_gb = float(os.environ.get("EXB_BENCH_C5_GB", "100"))
        R5 = int(total_gb * 1e9 / (total_bytes / (args.reads * world)) / world)  # the same on every rank
        if world == 1:
            buf5 = synth.gen_device(synth.gen_params("illumina", R5, seed=SEED, len_min=READ_LEN, len_max=READ_LEN), dev)
            shard5 = XD.Shard(buf5, 0, buf5.numel(), 0, True)
            job = XD.ShardedFastqTotals(shard5, None, ranges=[[0, buf5.numel(), 0]], rec_cap=R5 + 4096, max_lines=4 * R5 + 4096)

            def c5_step():
                job.scan()
                job.resolve_local(None, 0)
                job.finish_local()
            c5_total = job.agg
        else:
            shard5, gbuf5, s_prev5, own5 = build_shard(R5)
            job = XD.ShardedFastqTotals(shard5, grp, rec_cap=R5 + 4096, max_lines=4 * R5 + 4096)
            c5_step = job.step
            c5_total = job.total
        for _ in range(3):
            c5_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        c5_steps = max(3, min(args.steps, 5))
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(c5_steps):
            c5_step()
        c1.record()
        torch.cuda.synchronize()
        t5 = torch.tensor([c0.elapsed_time(c1) / c5_steps], dtype=torch.float64, device=dev)
        b5 = torch.tensor([shard5.hi - shard5.lo], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            dist.all_reduce(b5)
        got5 = [int(x) for x in c5_total.cpu().tolist()]
        if world == 1:
            blk5 = XD._result_block(job.ws).cpu().tolist()
            got5[6], got5[7] = blk5[0] & 3, int(blk5[2] != 0)
        XD.check_count(torch.tensor(got5))
        n5 = R5 * world
        assert got5[0] == n5 and got5[1] == READ_LEN * n5, (got5, n5)
        # cross-check of SUM(#GC): the scalar gc_content kernel's counts are not available per shard cut, so compare with an
        # independent whole-record scan of this rank's own records (exb_fastq_scan + exb_fastq_filter), summed over ranks
        own_view = buf5 if world == 1 else gbuf5[s_prev5:s_prev5 + own5]
        if world > 1 and (s_prev5 % 16) != 0:
            own_view = None  # misaligned view: skip the copy of tens of GB; ranks with an aligned view still check theirs
        chk = torch.zeros(2, dtype=torch.int64, device=dev)
        if own_view is not None:
            s5 = D.fastq_scan(own_view, _lib.F_SEQ, rec_cap=R5 + 4096)
            a5, _ = D.fastq_filter(s5, R5, [])
            chk[0], chk[1] = a5[2], 1
            del s5
        if world > 1:
            allchk = torch.zeros(2 * world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allchk, chk)
            allchk = allchk.view(world, 2).cpu().tolist()
        else:
            allchk = [chk.cpu().tolist()]
        if all(c[1] for c in allchk):
            assert sum(c[0] for c in allchk) == got5[2], (allchk, got5)
        ms5 = t5.item()
        job_fused = bool(job.fused)
        peak5, _ = measured_peak()
        c5 = {"workload": "C5: %.1f GB synthetic Illumina FASTQ, ONE file byte-range sharded over %d GPU(s), cuts inside records; "
                          "SELECT COUNT(*), SUM(#GC), SUM(length(sequence)), AVG(gc_content(sequence))" % (int(b5.item()) / 1e9, world),
              "total_bytes": int(b5.item()), "reads": n5, "ms_per_step": ms5, "value": int(b5.item()) / (ms5 * 1e-3) / 1e9, "unit": UNIT,
              "scaling": "strong", "steps": c5_steps, "frac_of_hbm_peak": int(b5.item()) / (ms5 * 1e-3) / 1e9 / (peak5 * world),
              "count": got5[0], "sum_len": got5[1], "sum_gc": got5[2], "avg_gc_content": got5[5] / 4294967296.0 / n5,
              "sum_gc_cross_checked": bool(all(c[1] for c in allchk)),
              "kernels": ("exb_fastq_scan_totals_begin (K1 with the sequence-line aggregates of all four phase hypotheses + line offsets) -> "
                          "block exchange -> exb_fastq_compose_prev -> exb_fastq_scan_totals_resolve (K2 picks the buckets) -> reduce")
                         if job_fused else
                         ("exb_fastq_scan_begin (K1 + line offsets) -> block exchange -> exb_fastq_compose_prev -> exb_fastq_scan_resolve "
                          "(EXB_F_SEQ | EXB_F_LOCAL_RECORDS) -> exb_fastq_seq_totals -> reduce")}
        del job

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = n_bytes / (scan_ms * 1e-3) / 1e9
        tr = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(world, args.reads),
            "reads_per_s": args.reads * world / (ms_step * 1e-3),
            "bytes_per_gpu": n_bytes, "records_passing": int(n_pass),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_per_launch(tr, n_bytes), "kernel": "fastq_tile_kernel<F_FUSED|F_QUAL> (TMA-fed byte pass; + offset scan + bucket-combine kernels, ~4% of the step)",
                         "algorithmic_bytes_per_launch": n_bytes, "kernel_ms": scan_ms, "peak_source": peak_src,
                         "note": "algorithmic bytes = input file bytes read once (SURVEY 8d); kernel_ms = CUDA events on the launching stream around one exb_fastq_scan_filter call (3 small memsets + tile kernel + 3 offset-scan kernels + combine kernel), i.e. an upper bound of the tile kernel's own duration"},
            "clocks": sampler.summary(),
            # N=1: fastq_tile_kernel, scan_reduce / scan_spine / scan_down (per-tile line offsets), fastq_fused_combine_kernel
            # (+ 3 cudaMemsetAsync); N>1: the combine kernel runs once, AFTER the exchange (fastq_final_state_kernel writes
            # the block that is exchanged), plus fastq_compose_prev_kernel (ranks >= 1) and, with the peer-memory
            # exchange, peer_allgather_kernel + peer_count_reduce_kernel
            # peer-memory exchange adds exb_peer_allgather_block + exb_peer_count_reduce (7 of this library's kernels per step)
            # launches of THIS library on rank 0 per step: tile + 3 scan + combine; N>1: + final_state (+ allgather + reduce
            # with the peer-memory exchange; rank 0 has no compose kernel)
            # N>1, single-exchange flavour (default): tile + 3 scan + candidates + peer_count_fused = 6 (NCCL: 5 + combine)
            "gpu_launches": (5 if world == 1 else 6) * args.steps,
        }
        if world > 1:
            line["exchange"] = exchange
        line["numa_node"] = numa_node
        if e2e:
            line["e2e"] = e2e
        if e2e_first:
            line["e2e_first_scan"] = e2e_first
        if e2e_pinned:
            line["e2e_pinned_image"] = e2e_pinned
        if path_rows:
            line["paths"] = path_rows
        if c5:
            line["c5"] = c5
        if cpu:
            line["cpu_baseline"] = cpu
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
