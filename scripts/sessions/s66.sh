#!/bin/bash
D=gpurun_out/s66; mkdir -p $D
timeout 500 ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate -s 3 -c 1 -f -o $D/inflate_v3 \
    python scripts/bench_paths.py --only bgzf --out $D/paths_ncu.json > $D/ncu.log 2>&1
echo "ncu exit $?"
