set -x
D=gpurun_out/${1:-s26}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_fasta.py tests/test_gpu_fullsize.py tests/test_gpu_reader2.py tests/test_gpu_writer.py tests/test_gpu_scalar_reader.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -5 $D/gputest.txt
python scripts/bench_paths.py --only c3 --out $D/paths.json 2>&1 | grep -v "^+" | tee $D/paths.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $D/paths_launches.csv \
    python scripts/bench_paths.py --only c3 --out $D/paths_under_ncu.json > $D/ncu_paths.log 2>&1
grep "fasta_tile_kernel<(bool)1>\|fasta_tile_kernel<1>" $D/paths_launches.csv | head -8
