#!/bin/bash
# Multi-GPU check on ONE box: bench.py under torchrun at N ranks (NCCL all-gather of the shard result blocks +
# all-reduce of the aggregates, exon_duckdb_b200/dist.py) and the reference arm beside it.
# usage: gpurun --gpus N --timeout 600 -- bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-mg}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; cat gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
