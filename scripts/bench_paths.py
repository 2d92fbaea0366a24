#!/usr/bin/env python
"""Device-resident throughput of every path of the scope table, one line each (tools/paths.py does the work;
bench.py reports the same rows under `paths`).

usage (GPU box): python scripts/bench_paths.py [--reads N] [--out gpurun_out/paths.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools import paths, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--ont-reads", type=int, default=40_000)
    ap.add_argument("--contigs", type=int, default=2400)
    ap.add_argument("--contig-len", type=int, default=500_000)
    ap.add_argument("--out", default="gpurun_out/paths.json")
    ap.add_argument("--only", default="c2,c4,c3,bgzf")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rep = paths.Report()
    only = args.only.split(",")
    if "c2" in only:
        buf = synth.gen_device(synth.gen_params("illumina", args.reads, seed=20), dev)
        paths.c2_paths(rep, buf, args.reads)
        del buf
        torch.cuda.empty_cache()
    if "c4" in only:
        paths.c4_paths(rep, dev, args.ont_reads)
    if "c3" in only:
        paths.c3_paths(rep, dev, args.contigs, args.contig_len)
    if "bgzf" in only:
        paths.bgzf_paths(rep, dev)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"peak_gbs": rep.peak, "rows": rep.rows}, f, indent=1)


if __name__ == "__main__":
    main()
