set -x
mkdir -p gpurun_out/s16
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s16/gputest.txt 2>&1
tail -8 gpurun_out/s16/gputest.txt
EXON_B200_TRACE=1 python bench.py --steps 10 --warmup 3 --no-paths --no-c5 --no-cpu > gpurun_out/s16/bench.json 2> gpurun_out/s16/bench.err
grep "exon_b200 reader" gpurun_out/s16/bench.err | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s16/bench.json').read().strip().splitlines()[-1])
print('e2e',{k:v for k,v in d.get('e2e',{}).items() if k not in ('note','api')})
print('pinned',d.get('e2e_pinned_image'))
PY
python scripts/bench_reader.py --out gpurun_out/s16/reader.json 2>&1 | tee gpurun_out/s16/reader.txt
python scripts/bench_duckdb.py --out gpurun_out/s16/duckdb.json 2>&1 | grep PRODUCT | tee gpurun_out/s16/duckdb.txt
