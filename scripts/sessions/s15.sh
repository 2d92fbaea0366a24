set -x
mkdir -p gpurun_out/s15
EXON_B200_TRACE=1 python bench.py --steps 5 --warmup 3 --no-paths --no-c5 --no-cpu > gpurun_out/s15/bench.json 2> gpurun_out/s15/bench.err
grep "exon_b200 reader" gpurun_out/s15/bench.err | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s15/bench.json').read().strip().splitlines()[-1])
print('e2e',{k:v for k,v in d.get('e2e',{}).items() if k not in ('note','api')})
PY
python - <<'PY'
import sys, time, ctypes as C, os
sys.path.insert(0,'.')
from tools import synth
from exon_duckdb_b200 import _lib
from exon_duckdb_b200._lib import lib, check
path='/dev/shm/seven.fastq'
synth.gen_host(synth.gen_params("illumina", 20_000_000, seed=20)).tofile(path)
sz=os.path.getsize(path)
def run(filt):
    h=C.c_void_p(); t0=time.perf_counter()
    o=_lib.reader_options(column_mask=0)
    check(lib().exb_reader_open2(path.encode(), b"fastq", None, 2048, filt, C.byref(o), C.byref(h)))
    n=C.c_int64(); check(lib().exb_reader_count(h, C.byref(n)))
    lib().exb_reader_close(h)
    return time.perf_counter()-t0, n.value
for filt in (None, b"mean_quality(quality_scores)>30"):
    ts=[run(filt)[0] for _ in range(6)]
    print(filt, "best %.1f ms %.2f GB/s all %s"%(min(ts)*1e3, sz/1e9/min(ts), " ".join("%.0f"%(t*1e3) for t in ts)), flush=True)
import torch
x=torch.zeros(1<<30, dtype=torch.uint8, device='cuda')
for filt in (None, b"mean_quality(quality_scores)>30"):
    ts=[run(filt)[0] for _ in range(6)]
    print("after torch cuda init", filt, "best %.1f ms %.2f GB/s all %s"%(min(ts)*1e3, sz/1e9/min(ts), " ".join("%.0f"%(t*1e3) for t in ts)), flush=True)
PY
