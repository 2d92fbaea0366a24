#!/bin/bash
# Per-kernel launch list of every path of scripts/bench_paths.py + the native reader's end-to-end throughput.
# usage: gpurun --timeout 900 -- bash scripts/gpu_paths_profile.sh <tag>
TAG=${1:-pp}
mkdir -p gpurun_out
timeout 300 python scripts/bench_reader.py --out gpurun_out/${TAG}_reader.json > gpurun_out/${TAG}_reader.log 2>&1
echo "reader exit $?"; tail -12 gpurun_out/${TAG}_reader.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_paths_launches.csv \
    python scripts/bench_paths.py --iters 1 --warm 1 --out gpurun_out/${TAG}_paths_ncu.json > gpurun_out/${TAG}_paths_ncu.log 2>&1
echo "ncu paths exit $?"; tail -3 gpurun_out/${TAG}_paths_ncu.log
