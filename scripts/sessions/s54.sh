#!/bin/bash
# inflate kernel A/B: (A) tree as is; (B) 9-bit literal table, 6 CTAs / SM -- rebuilt on the box
D=gpurun_out/s54; mkdir -p $D
timeout 300 python scripts/bench_paths.py --only bgzf --out $D/paths_a.json 2>&1 | tail -2
cd exon_duckdb_b200/csrc
for V in "-DIFL_LB=9 -DIFL_DB=8 -DIFL_MINB=6" "-DIFL_LB=9 -DIFL_DB=7 -DIFL_MINB=8"; do
  rm -f inflate.o; make -s NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $V" 2>&1 | tail -2
  grep "Used" inflate.ptxas.log
  (cd ../..; echo "variant $V"; timeout 300 python scripts/bench_paths.py --only bgzf --out $D/paths_b.json 2>&1 | tail -2; timeout 300 python -m pytest tests/test_inflate.py -m gpu -x -q 2>&1 | tail -1)
done
