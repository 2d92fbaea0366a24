#!/bin/bash
# One gpurun call for the round's evidence set (1 GPU): GPU parity tests, smoke, the driver-contract bench lines (own arm and
# reference arm), ncu launch list of the bench, full captures of K1 and of the FASTA compaction kernel, every other path,
# the reader C ABI and SQL through DuckDB.
# usage: gpurun --timeout 2400 -- bash scripts/gpu_final.sh <outdir under gpurun_out>
set -x
D=gpurun_out/${1:-final}
mkdir -p $D
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $D/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $D/gputest.txt 2>&1
echo "pytest exit $?" | tee -a $D/gputest.txt
tail -4 $D/gputest.txt
python __graft_entry__.py smoke > $D/smoke.txt 2>&1; tail -1 $D/smoke.txt
timeout 900 python bench.py > $D/bench.json 2> $D/bench.err
echo "bench exit $?"; tail -2 $D/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_ref.json 2>> $D/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-c5 --no-paths > $D/ncu_launches.log 2>&1
EXB_BENCH_READS=4000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastq_tile_kernel -s 3 -c 1 \
    -f -o $D/fastq_scan python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-c5 --no-paths > $D/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fasta_tile_kernel -s 15 -c 1 -f -o $D/fasta_k3 \
    python scripts/bench_paths.py --only c3 --out $D/paths_under_ncu.json > $D/ncu_k3.log 2>&1
python scripts/bench_paths.py --out $D/paths.json 2>&1 | grep -v "^+" > $D/paths.txt
python scripts/bench_reader.py --out $D/reader.json > $D/reader.txt 2>&1
python scripts/bench_duckdb.py --out $D/duckdb.json > $D/duckdb.txt 2>&1
tail -3 $D/duckdb.txt
