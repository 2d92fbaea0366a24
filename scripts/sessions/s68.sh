#!/bin/bash
timeout 600 python -m pytest tests/test_duckdb_ext.py tests/test_gpu_reader2.py -m gpu -x -q 2>&1 | tail -12
