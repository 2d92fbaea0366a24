set -x
mkdir -p gpurun_out/s6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s6/gputest.txt 2>&1
tail -5 gpurun_out/s6/gputest.txt
EXON_B200_TRACE=1 python scripts/bench_reader.py --out gpurun_out/s6/reader.json > gpurun_out/s6/reader.txt 2>&1
grep -v "^exon_b200 reader" gpurun_out/s6/reader.txt
grep "^exon_b200 reader" gpurun_out/s6/reader.txt | sed -n '3p;13p;28p'
python bench.py --steps 5 --warmup 3 > gpurun_out/s6/bench.json 2> gpurun_out/s6/bench.err
tail -3 gpurun_out/s6/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s6/bench.json'))
print({k:d[k] for k in ('value','ms_per_step')})
print('e2e',d.get('e2e'))
print('pinned',d.get('e2e_pinned_image'))
for r in d.get('paths',[]): print("%-70s %8.3f ms %8.1f GB/s %.3f"%(r['path'],r['ms_median'],r['GB/s'],r['frac']))
print(d.get('cpu_baseline'))
PY
