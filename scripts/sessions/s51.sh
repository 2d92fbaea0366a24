#!/bin/bash
# BGZF inflate kernel: A/B of pended match stores + one full ncu capture of each
D=gpurun_out/s51; mkdir -p $D
for P in 0 1; do
  EXB_INFLATE_PEND=$P timeout 300 python scripts/bench_paths.py --only bgzf --out $D/paths_pend$P.json 2>&1 | tail -2
  EXB_INFLATE_PEND=$P timeout 500 ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate -s 3 -c 1 -f -o $D/inflate_pend$P \
      python scripts/bench_paths.py --only bgzf --out $D/paths_ncu.json > $D/ncu_pend$P.log 2>&1
  echo "ncu exit $?"
done
ls -la $D
