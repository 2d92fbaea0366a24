#!/bin/bash
# A/B of two builds of the library on the same box: usage gpurun -- bash scripts/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
for v in A B A B; do
  if [ $v = B ]; then export EXON_B200_LIB=$PWD/exon_duckdb_b200/libexon_b200_B.so; else unset EXON_B200_LIB; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_$v.csv python scripts/profile_table.py > /dev/null 2>&1
  echo "variant $v:"; grep gather_span gpurun_out/${TAG}_$v.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -12 | tr '\n' ' '; echo
done
