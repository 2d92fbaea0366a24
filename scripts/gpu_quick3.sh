#!/bin/bash
# usage: gpurun --timeout 1200 -- bash scripts/gpu_quick3.sh <tag> <pytest -k expr> <kernel regex> <skip>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
EXON_B200_TRACE=1 timeout 300 python scripts/bench_reader.py --repeat 2 --out gpurun_out/${TAG}_reader.json > gpurun_out/${TAG}_reader.log 2>&1
echo "reader exit $?"; tail -28 gpurun_out/${TAG}_reader.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$3 -s ${4:-0} -c 1 -f -o gpurun_out/${TAG}_cap \
    python scripts/profile_table.py > gpurun_out/${TAG}_cap.log 2>&1
echo "ncu exit $?"; python scripts/profile_table.py 2>&1 | tail -4
