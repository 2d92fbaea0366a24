#!/bin/bash
# Quick iteration on the GPU: FASTQ parity tests, a short bench, one full ncu capture of the scan kernel.
# usage: gpurun --timeout 900 -- bash scripts/gpu_iter.sh <tag> [pytest -k expr]
TAG=${1:-it}
KEXPR=${2:-fastq}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastq_tile_kernel -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_fastq_scan python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
