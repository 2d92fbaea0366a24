set -x
D=gpurun_out/${1:-s27}
mkdir -p $D
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fasta_tile_kernel -s 15 -c 1 -f -o $D/fasta_k3 \
    python scripts/bench_paths.py --only c3 --out $D/paths_under_ncu.json > $D/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 $D/ncu_full.log
