#!/bin/bash
# final evidence of the round on the final tree (1 GPU): GPU parity suite, smoke, driver-contract bench lines, paths, reader
D=gpurun_out/s60; mkdir -p $D
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $D/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $D/gputest.txt 2>&1; echo "pytest exit $?" | tee -a $D/gputest.txt; tail -3 $D/gputest.txt
python __graft_entry__.py smoke > $D/smoke.txt 2>&1; tail -1 $D/smoke.txt
SECONDS=0
timeout 900 python bench.py > $D/bench.json 2> $D/bench.err; echo "bench exit $? in $SECONDS s"; tail -2 $D/bench.err
SECONDS=0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_ref.json 2>> $D/bench.err; echo "ref exit $? in $SECONDS s"
python scripts/bench_paths.py --out $D/paths.json 2>&1 | grep -v "^+" > $D/paths.txt; tail -3 $D/paths.txt
python scripts/bench_reader.py --out $D/reader.json > $D/reader.txt 2>&1; tail -6 $D/reader.txt
