set -x
D=gpurun_out/${1:-s33}
mkdir -p $D
# two independent single-GPU runs at the same time
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 30 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > $D/solo0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 30 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > $D/solo1.json 2>/dev/null &
wait
for f in solo0 solo1; do python -c "
import json; d=json.loads(open('$D/$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['clocks'])"; done
for MODE in 1 0; do
EXB_EXCHANGE_FUSED=$MODE EXB_BENCH_STAGES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$MODE bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-c5 > $D/n2_fused$MODE.json 2> $D/n2_fused$MODE.err
python -c "
import json; d=json.loads(open('$D/n2_fused$MODE.json').read().strip().splitlines()[-1]); print('fused$MODE', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('stage_ms'), d['clocks'])"
done
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 30 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > $D/gpu1_alone.json 2>/dev/null
python -c "
import json; d=json.loads(open('$D/gpu1_alone.json').read().strip().splitlines()[-1]); print('gpu1 alone', d['value'], d['ms_per_step'])"
