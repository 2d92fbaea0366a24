set -x
D=gpurun_out/${1:-s44}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_fullsize.py tests/test_gpu_scalar_reader.py tests/test_gpu_reader2.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -4 $D/gputest.txt
python scripts/bench_paths.py --only c4,c2 --out $D/paths.json 2>&1 | grep -v "^+" | tee $D/paths.txt
