#!/bin/bash
# BGZF device inflate: parity tests, kernel throughput, reader end to end
D=gpurun_out/s50; mkdir -p $D
timeout 600 python -m pytest tests/test_inflate.py -m gpu -x -q > $D/pytest.txt 2>&1; echo "pytest exit $?"; tail -15 $D/pytest.txt
timeout 300 python scripts/bench_paths.py --only bgzf --out $D/paths.json 2>&1 | tail -5
timeout 600 python scripts/bench_reader.py --cases bgzf --out $D/reader.json > $D/reader.txt 2>&1; cat $D/reader.txt | tail -8
