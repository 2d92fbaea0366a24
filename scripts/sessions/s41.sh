D=gpurun_out/${1:-s41}
mkdir -p $D
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-c5 > $D/bench_n2.json 2> $D/bench_n2.err &
for i in $(seq 1 14); do sleep 3; echo "--- t=$((3*i))s"; nvidia-smi --query-compute-apps=pid,gpu_bus_id,used_memory --format=csv,noheader; nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv,noheader; done > $D/apps.txt 2>&1
wait
cat $D/apps.txt | tail -40
python -c "
import json; d=json.loads(open('$D/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['numa_node'], d['e2e']['value'])"
