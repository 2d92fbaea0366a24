#!/bin/bash
# usage: gpurun --timeout 600 -- bash scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/${TAG}_pytest.log
fi
python scripts/profile_table.py 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_table_launches.csv python scripts/profile_table.py > gpurun_out/${TAG}_table_ncu.log 2>&1
echo "ncu exit $?"
