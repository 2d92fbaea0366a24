set -x
D=gpurun_out/${1:-s25}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_fullsize.py tests/test_gpu_reader2.py tests/test_gpu_dist.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -5 $D/gputest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/paths_launches.csv \
    python scripts/bench_paths.py --only c2,c4 --out $D/paths_under_ncu.json > $D/ncu_paths.log 2>&1
python scripts/bench_paths.py --only c2,c4 --out $D/paths.json 2>&1 | grep -v "^+" | tee $D/paths.txt
