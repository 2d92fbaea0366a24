set -x
mkdir -p gpurun_out/s8
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_fullsize.py tests/test_gpu_reader2.py tests/test_gpu_scalar_reader.py -m gpu -x -q > gpurun_out/s8/gputest.txt 2>&1
tail -5 gpurun_out/s8/gputest.txt
python scripts/bench_paths.py --only c2 --out gpurun_out/s8/paths.json 2>&1 | tee gpurun_out/s8/paths.txt
EXON_B200_TRACE=1 python scripts/bench_reader.py --out gpurun_out/s8/reader.json > gpurun_out/s8/reader.txt 2>&1
grep -v "^exon_b200 reader" gpurun_out/s8/reader.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fastq_split_kernel -c 1 -o gpurun_out/s8/split -f python scripts/bench_paths.py --only c2 --out /tmp/x.json > gpurun_out/s8/ncu.log 2>&1
tail -3 gpurun_out/s8/ncu.log
