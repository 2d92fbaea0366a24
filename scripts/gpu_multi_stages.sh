#!/bin/bash
# usage: gpurun --gpus N --timeout 600 -- bash scripts/gpu_multi_stages.sh <tag> <N>
TAG=${1:-mg}
N=${2:-8}
mkdir -p gpurun_out
EXB_BENCH_STAGES=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/${TAG}_stages_n$N.json 2> gpurun_out/${TAG}_stages_n$N.err
echo "exit $?"; grep "stages" gpurun_out/${TAG}_stages_n$N.err; cut -c1-200 gpurun_out/${TAG}_stages_n$N.json
nproc; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
