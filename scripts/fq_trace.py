#!/usr/bin/env python
"""Per-tile timeline of the FASTQ scan's chain (debug build: make -C exon_duckdb_b200/csrc trace).

usage (GPU box): EXON_B200_LIB=exon_duckdb_b200/libexon_b200_trace.so python scripts/fq_trace.py [reads] [out.npy]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from exon_duckdb_b200 import _lib, device as D

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/fq_trace.npy"
dev = torch.device("cuda:0")
L = _lib.lib()
p = _lib.gen_params("illumina", reads, seed=20)
buf = D.gen_device(p, dev)
preds = [("mean_quality", ">", 30.0)]
c = D.fastq_scan_filter(buf, preds)
c.validate()
for _ in range(3):
    D.fastq_scan_filter(buf, preds, out=c)
torch.cuda.synchronize()
n_tiles = (buf.numel() + 1 + 2047) // 2048
trace = torch.zeros(n_tiles * 4, dtype=torch.int64, device=dev)
L.exb_debug_set_fq_trace.argtypes = [C.c_void_p]
assert L.exb_debug_set_fq_trace(trace.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
D.fastq_scan_filter(buf, preds, out=c)
e1.record()
torch.cuda.synchronize()
print("traced launch: %.3f ms for %d bytes (%d tiles)" % (e0.elapsed_time(e1), buf.numel(), n_tiles))
t = trace.cpu().numpy().reshape(n_tiles, 4)
np.save(out, t)
t0 = t[:, 3].min()
pub, res, tick = t[:, 0] - t0, t[:, 1] - t0, t[:, 3] - t0
spins, last = t[:, 2] & 0xFFFFFFFF, (t[:, 2] >> 32) & 1
print("span %.1f us" % (res.max() / 1e3))
print("publish-ticket   (stage A)  : median %.2f us  p90 %.2f  max %.2f" % tuple(np.percentile(pub - tick, [50, 90, 100]) / 1e3))
print("resolved-publish (chain)    : median %.2f us  p90 %.2f  max %.2f" % tuple(np.percentile(res - pub, [50, 90, 100]) / 1e3))
print("spins: median %d p90 %d max %d ; last arrivers %d" % (np.median(spins), np.percentile(spins, 90), spins.max(), last.sum()))
for k in range(0, min(n_tiles, 40000), 4000):
    print("tile %6d: ticket %8.2f publish %8.2f resolved %8.2f us spins %d last %d" % (k, tick[k] / 1e3, pub[k] / 1e3, res[k] / 1e3, spins[k], last[k]))
