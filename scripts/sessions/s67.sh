#!/bin/bash
# inflate kernel: 32-bit word index in the bit reader (A), and the same at 5 CTAs / SM = 48 registers (B)
timeout 300 python -m pytest tests/test_inflate.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python scripts/bench_paths.py --only bgzf --out gpurun_out/s67_a.json 2>&1 | tail -2
cd exon_duckdb_b200/csrc
rm -f inflate.o; make -s NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v -DIFL_MINB=5" 2>&1 | tail -2
grep "Used" inflate.ptxas.log
cd ../..; timeout 300 python scripts/bench_paths.py --only bgzf --out gpurun_out/s67_b.json 2>&1 | tail -2
