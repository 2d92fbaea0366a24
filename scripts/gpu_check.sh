#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, one full capture of the scan kernel.
# usage: gpurun --timeout 1700 -- bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
EXB_BENCH_READS=4000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastq_tile_kernel -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_fastq_scan python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 python scripts/bench_paths.py --out gpurun_out/${TAG}_paths.json > gpurun_out/${TAG}_paths.log 2>&1
echo "paths exit $?"; tail -30 gpurun_out/${TAG}_paths.log
timeout 300 python scripts/bench_reader.py --out gpurun_out/${TAG}_reader.json > gpurun_out/${TAG}_reader.log 2>&1
echo "reader exit $?"; tail -9 gpurun_out/${TAG}_reader.log
