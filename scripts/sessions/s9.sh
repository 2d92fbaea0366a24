set -x
mkdir -p gpurun_out/s9
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/s9/gputest.txt 2>&1
tail -5 gpurun_out/s9/gputest.txt
python scripts/bench_paths.py --only c2,c4 --out gpurun_out/s9/paths.json 2>&1 | tee gpurun_out/s9/paths.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fastq_split_kernel -c 1 -o gpurun_out/s9/split -f python scripts/bench_paths.py --only c2 --out /tmp/x.json > gpurun_out/s9/ncu.log 2>&1
tail -3 gpurun_out/s9/ncu.log
