#!/usr/bin/env python
"""What a sequence of back-to-back scans of a FRESH file costs while its page cache is being registered in the background
(reader.cu FileMap::start_register): scan 1 takes the copy path and triggers the registration, scan 2 may run while
cudaHostRegister holds its locks, later scans DMA from the registered mapping.  Prints wall time and I/O path per scan."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exon_duckdb_b200 import _lib
from tools import synth

L = _lib.lib()
path = "/dev/shm/exb_probe_seq.fastq"
synth.gen_host(synth.gen_params("illumina", 8_000_000, seed=21)).tofile(path)
size = os.path.getsize(path)
filt = b"mean_quality(quality_scores)>30.0"


def scan(flags=0):
    o = _lib.reader_options(column_mask=0, flags=flags)
    h = C.c_void_p()
    t0 = time.perf_counter()
    _lib.check(L.exb_reader_open2(path.encode(), b"fastq", None, 2048, filt, C.byref(o), C.byref(h)))
    n = C.c_int64()
    _lib.check(L.exb_reader_count(h, C.byref(n)))
    direct = L.exb_reader_io_path(h)
    L.exb_reader_close(h)
    return time.perf_counter() - t0, direct


# warm the process (context, pools) on the copy path without triggering the registration
for _ in range(2):
    scan(_lib.RD_COPY_IO)
print("copy path, warm:            %.1f ms" % (scan(_lib.RD_COPY_IO)[0] * 1e3))
def show(i, dt, direct):
    print("scan %d: %7.1f ms  %5.1f GB/s  path=%s  cache_state_after=%d" % (i, dt * 1e3, size / dt / 1e9, "registered" if direct else "copy", L.exb_file_cache_state(path.encode())), flush=True)


for i in range(4):  # back to back: the registration does not get its turn
    show(i + 1, *scan())
t0 = time.perf_counter()
while L.exb_file_cache_state(path.encode()) == 1 and time.perf_counter() - t0 < 30:
    time.sleep(0.005)
print("pause: registered after %.0f ms of idle time" % ((time.perf_counter() - t0) * 1e3))
for i in range(4, 7):
    show(i + 1, *scan())
os.unlink(path)
