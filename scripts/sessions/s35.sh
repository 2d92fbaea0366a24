set -x
D=gpurun_out/${1:-s35}
mkdir -p $D
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-c5 > $D/bench_n$N.json 2> $D/bench_n$N.err
python -c "
import json; d=json.loads(open('$D/bench_n$N.json').read().strip().splitlines()[-1]); print('n$N', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --steps 20 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > $D/bench_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('$D/bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
