set -x
D=gpurun_out/${1:-s31}
mkdir -p $D
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $D/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $D/bench_n$N.json 2> $D/bench_n$N.err
tail -3 $D/bench_n$N.err
python - <<PY
import json
d=json.loads(open('$D/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','exchange','numa_node')}); print(d.get('e2e')); print({k:v for k,v in (d.get('c5') or {}).items() if k not in ('kernels','workload')})
PY
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > $D/bench_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('$D/bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
