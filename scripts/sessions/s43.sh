set -x
D=gpurun_out/${1:-s43}
mkdir -p $D
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastq_split_kernel -s 2 -c 1 -f -o $D/split_map \
    python scripts/bench_paths.py --only c4 --out $D/paths_under_ncu.json > $D/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 $D/ncu_full.log
