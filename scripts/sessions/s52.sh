#!/bin/bash
# full GPU suite after the fused-loop rewrite + BGZF through the reader and DuckDB
D=gpurun_out/s52; mkdir -p $D
timeout 1500 python -m pytest tests -m gpu -x -q > $D/gputest.txt 2>&1; echo "pytest exit $?"; tail -5 $D/gputest.txt
timeout 600 python scripts/bench_reader.py --cases bgzf --out $D/reader.json > $D/reader.txt 2>&1; tail -5 $D/reader.txt
timeout 900 python scripts/bench_duckdb.py --out $D/duckdb.json > $D/duckdb.txt 2>&1; grep -i "bgzf\|FAILED\|ERROR" $D/duckdb.txt
