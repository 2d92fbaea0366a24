set -x
mkdir -p gpurun_out/s7
python - <<'PY'
import sys
sys.path.insert(0,'.')
from tools import synth
synth.gen_host(synth.gen_params("illumina", 8_000_000, seed=20)).tofile('/dev/shm/exb_bench.fastq')
PY
IOB_QUICK=1 ./build/rt/iobench2 /dev/shm/exb_bench.fastq 2>&1 | tee gpurun_out/s7/calib.txt
cat > /tmp/rd.py <<'PY'
import sys, time, ctypes as C, os
sys.path.insert(0,'.')
from exon_duckdb_b200 import _lib
from exon_duckdb_b200._lib import lib, check
path='/dev/shm/exb_bench.fastq'
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
    return time.perf_counter()-t0, rows
for name,mask,filt,count in [("count",0,None,True),("4col",15,None,False)]:
    ts=[]
    for rep in range(6):
        dt,rows=run(mask,filt,count); ts.append(dt)
    print("%-8s best %.1f ms %.2f GB/s | median %.1f ms | all %s"%(name,min(ts)*1e3,sz/1e9/min(ts),sorted(ts)[3]*1e3," ".join("%.0f"%(t*1e3) for t in ts)),flush=True)
PY
for V in "X=1" "EXON_B200_NO_AVX512=1" "EXON_B200_IO_THREADS=12" "EXON_B200_IO_THREADS=16" "EXON_B200_IO_THREADS=4" "EXON_B200_NO_AVX512=1 EXON_B200_IO_THREADS=16" "X=2"; do echo "== $V"; env $V EXON_B200_TRACE=1 python /tmp/rd.py 2>&1 | grep -v "^exon_b200 reader" ; done > gpurun_out/s7/ab.txt 2>&1
cat gpurun_out/s7/ab.txt
IOB_QUICK=1 ./build/rt/iobench2 /dev/shm/exb_bench.fastq 2>&1 | tee -a gpurun_out/s7/calib.txt
EXON_B200_TRACE=2 python /tmp/rd.py 2> gpurun_out/s7/trace2.txt | tail -3
grep chunk gpurun_out/s7/trace2.txt | sed -n '60,70p;300,310p'
timeout 600 python -m pytest tests/test_duckdb_ext.py -m gpu -x -q 2>&1 | tail -3
