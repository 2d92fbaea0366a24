#!/bin/bash
timeout 300 python -m pytest tests/test_inflate.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python scripts/bench_paths.py --only bgzf --out gpurun_out/s63_paths.json 2>&1 | tail -2
