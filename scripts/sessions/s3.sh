set -x
mkdir -p gpurun_out/s3
python - <<'PY'
import sys
sys.path.insert(0,'.')
from tools import synth
synth.gen_host(synth.gen_params("illumina", 8_000_000, seed=20)).tofile('/dev/shm/big.fastq')
PY
cat > /tmp/rd.py <<'PY'
import sys, time, ctypes as C
sys.path.insert(0,'.')
from exon_duckdb_b200 import _lib
from exon_duckdb_b200._lib import lib, check
import os
path='/dev/shm/big.fastq'
sz=os.path.getsize(path)
def run(mask, filt, count):
    h=C.c_void_p()
    t0=time.perf_counter()
    o=_lib.reader_options(column_mask=mask, flags=_lib.RD_STRING_T|_lib.RD_NO_OFFSETS)
    check(lib().exb_reader_open2(path.encode(), b"fastq", None, 2048, filt, C.byref(o), C.byref(h)))
    rows=0
    if count:
        n=C.c_int64(); check(lib().exb_reader_count(h, C.byref(n))); rows=n.value
    else:
        b=_lib.Batch()
        while True:
            check(lib().exb_reader_next(h, C.byref(b)))
            if b.n_rows==0: break
            rows+=b.n_rows
            lib().exb_batch_release(C.byref(b))
    lib().exb_reader_close(h)
    dt=time.perf_counter()-t0
    return dt, rows
for name,mask,filt,count in [("count",0,None,True),("count+filter",0,b"mean_quality(quality_scores)>30",True),("4col",15,None,False),("seq",4,None,False)]:
    for rep in range(3):
        dt,rows=run(mask,filt,count)
        print("%-14s %.1f ms %.2f GB/s rows %d"%(name,dt*1e3,sz/1e9/dt,rows),flush=True)
PY
for T in 4 6 8 12 16; do echo "== IO threads $T"; EXON_B200_IO_THREADS=$T EXON_B200_TRACE=1 python /tmp/rd.py 2>&1 | grep -v "^exon_b200 reader" ; done > gpurun_out/s3/threads.txt 2>&1
cat gpurun_out/s3/threads.txt
EXON_B200_TRACE=2 python /tmp/rd.py > gpurun_out/s3/trace2.txt 2>&1
grep -c chunk gpurun_out/s3/trace2.txt
sed -n 1,30p gpurun_out/s3/trace2.txt
