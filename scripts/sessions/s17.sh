set -x
mkdir -p gpurun_out/s17
timeout 900 python -m pytest tests/test_duckdb_ext.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/s17/gputest.txt 2>&1
tail -4 gpurun_out/s17/gputest.txt
python scripts/bench_duckdb_multigpu.py --out gpurun_out/s17/duckdb_mg.json 2>&1 | grep -v "^+" | tee gpurun_out/s17/duckdb_mg.txt
for MODE in 1 0; do
EXB_EXCHANGE_FUSED=$MODE python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$MODE bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-c5 > gpurun_out/s17/bench_n2_fused$MODE.json 2> gpurun_out/s17/bench_n2_fused$MODE.err
done
python bench.py --steps 20 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > gpurun_out/s17/bench_n1.json 2>/dev/null
python - <<'PY'
import json
for f in ('bench_n1','bench_n2_fused1','bench_n2_fused0'):
    try:
        d=json.loads(open('gpurun_out/s17/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'])
    except Exception as e: print(f, 'failed', e)
PY
bash scripts/gpu_sanitize.sh gpurun_out/s17/sanitize > /dev/null 2>&1; cat gpurun_out/s17/sanitize/summary.txt
