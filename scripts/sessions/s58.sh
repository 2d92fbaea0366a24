#!/bin/bash
D=gpurun_out/s58; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_reader2.py -m gpu -x -q -k registered 2>&1 | tail -2
python scripts/probe_register_sequence.py 2>&1 | tail -10 | tee $D/register_sequence.txt
python scripts/probe_register_sequence.py 2>&1 | tail -10 | tee -a $D/register_sequence.txt
timeout 900 python scripts/bench_duckdb.py --out $D/duckdb.json > $D/duckdb.txt 2>&1; grep "PRODUCT" $D/duckdb.txt | cut -c1-170
