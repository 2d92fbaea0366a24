set -x
mkdir -p gpurun_out/s13
python scripts/bench_duckdb.py --out gpurun_out/s13/duckdb.json 2>&1 | tee gpurun_out/s13/duckdb.txt
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/s13/bench.json 2> gpurun_out/s13/bench.err
tail -5 gpurun_out/s13/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s13/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
print('e2e',{k:v for k,v in d.get('e2e',{}).items() if k not in ('note','api')})
print('pinned',d.get('e2e_pinned_image'))
for r in d.get('paths',[]): print("%-86s %8.3f ms %8.1f GB/s %.3f"%(r['path'],r['ms_median'],r['GB/s'],r['frac']))
print('c5',{k:v for k,v in (d.get('c5') or {}).items() if k not in ('kernels','workload')})
print(d.get('cpu_baseline'))
PY
