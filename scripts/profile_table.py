#!/usr/bin/env python
"""Run the table-building paths once each (after a warm-up) so that an ncu launch list shows their kernels in order.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/profile_table.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from exon_duckdb_b200 import _lib, device as D
from tools import synth

dev = torch.device("cuda:0")
reads = int(os.environ.get("EXB_READS", 4_000_000))
buf = synth.gen_device(synth.gen_params("illumina", reads, seed=20), dev)
preds = [("mean_quality", ">", 30.0)]
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tab = D.fastq_table(buf)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    tab2 = D.fastq_table(buf, columns=["name", "sequence"], preds=preds)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("fastq_table all: %.3f ms   filtered name+sequence: %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
del tab, tab2, buf
torch.cuda.empty_cache()
fa = synth.gen_device(synth.gen_params("fasta", 2400, seed=3, len_min=500000, len_max=500000, wrap=60), dev)
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fs = D.fasta_scan(fa, compact=False)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    fs2 = D.fasta_scan(fa, compact=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("fasta_scan: %.3f ms   with compaction: %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
