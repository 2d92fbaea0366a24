#!/usr/bin/env python
"""Rewrite a gzip (or plain) FASTA / FASTQ file as BGZF, the blocked gzip layout the reader inflates on the device
(DESIGN 6b): single-stream .gz files keep the streaming host decoder, BGZF ones are decompressed by the GPU.

    python -m tools.bgzip in.fastq.gz out.fastq.gz [--threads 16] [--level 6]

The output is a valid gzip file for every other tool.  Host-side only: zlib through exb_bgzf_compress_host."""
import argparse
import ctypes as C
import gzip
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exon_duckdb_b200 import _lib

PIECE = 65280 * 1024  # text bytes compressed per call (a whole number of members)


def recompress(src, dst, threads=16, level=6):
    L = _lib.lib()
    with open(src, "rb") as f:
        magic = f.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    cap = L.exb_bgzf_compress_bound(PIECE)
    out = C.create_string_buffer(cap)
    n_out = C.c_int64()
    text_bytes = file_bytes = 0
    with opener(src, "rb") as fin, open(dst, "wb") as fout:
        while True:
            piece = fin.read(PIECE)
            last = len(piece) < PIECE
            _lib.check(L.exb_bgzf_compress_host(piece, len(piece), level, threads, 1 if last else 0, out, cap, C.byref(n_out)))
            fout.write(out.raw[:n_out.value])
            text_bytes += len(piece)
            file_bytes += n_out.value
            if last:
                break
    return text_bytes, file_bytes


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--threads", type=int, default=min(16, os.cpu_count() or 4))
    ap.add_argument("--level", type=int, default=6)
    a = ap.parse_args()
    t, f = recompress(a.src, a.dst, a.threads, a.level)
    print("%s: %d text bytes -> %d bytes of BGZF (%s)" % (a.dst, t, f, "ratio %.2f" % (t / f) if f else "empty"))


if __name__ == "__main__":
    main()
