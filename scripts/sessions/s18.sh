set -x
mkdir -p gpurun_out/s18
N=$(nvidia-smi -L | wc -l)
for MODE in 1 0; do
EXB_EXCHANGE_FUSED=$MODE python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$MODE bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-c5 > gpurun_out/s18/bench_n${N}_fused$MODE.json 2> gpurun_out/s18/bench_n${N}_fused$MODE.err
done
python bench.py --steps 20 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > gpurun_out/s18/bench_n1.json 2>/dev/null
python - <<PY
import json
for f in ('bench_n1','bench_n${N}_fused1','bench_n${N}_fused0'):
    try:
        d=json.loads(open('gpurun_out/s18/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'])
    except Exception as e: print(f, 'failed', e)
PY
