#!/bin/bash
# One multi-GPU box: N-rank bench (peer-memory exchange), then the single-GPU parity tests and per-path throughput.
# usage: gpurun --gpus N --timeout 1500 -- bash scripts/gpu_round.sh <tag> <N>
TAG=${1:-rr}
N=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_n$N.err | tail -8
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 400 python scripts/bench_paths.py --out gpurun_out/${TAG}_paths.json > gpurun_out/${TAG}_paths.log 2>&1
echo "paths exit $?"; tail -16 gpurun_out/${TAG}_paths.log
