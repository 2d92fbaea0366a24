#!/bin/bash
# Per-launch duration + DRAM bytes of every kernel of the table / FASTA paths (scripts/profile_table.py).
# usage: gpurun --timeout 900 -- bash scripts/gpu_traffic.sh <tag>
TAG=${1:-tr}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_traffic.csv python scripts/profile_table.py > gpurun_out/${TAG}_traffic.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${TAG}_traffic.log
