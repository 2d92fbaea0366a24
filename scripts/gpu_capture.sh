#!/bin/bash
# One full ncu capture of one kernel of scripts/profile_table.py.
# usage: gpurun --timeout 600 -- bash scripts/gpu_capture.sh <tag> <kernel-regex> <skip>
TAG=${1:-cap}
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${3:-0} -c 1 -f -o gpurun_out/${TAG} \
    python scripts/profile_table.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
python scripts/profile_table.py 2>&1 | tail -4
