#!/bin/bash
# usage: gpurun --timeout 1200 -- bash scripts/gpu_quick2.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/bench_reader.py --out gpurun_out/${TAG}_reader.json > gpurun_out/${TAG}_reader.log 2>&1
echo "reader exit $?"; tail -12 gpurun_out/${TAG}_reader.log
python scripts/profile_table.py 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_table_launches.csv python scripts/profile_table.py > gpurun_out/${TAG}_table_ncu.log 2>&1
echo "ncu exit $?"
