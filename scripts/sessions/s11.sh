set -x
mkdir -p gpurun_out/s11
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s11/gputest.txt 2>&1
tail -8 gpurun_out/s11/gputest.txt
python scripts/bench_paths.py --only c2,c4 --out gpurun_out/s11/paths.json 2>&1 | tee gpurun_out/s11/paths.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fastq_split_kernel -c 1 -o gpurun_out/s11/split -f python scripts/bench_paths.py --only c2 --out /tmp/x.json > gpurun_out/s11/ncu.log 2>&1
tail -3 gpurun_out/s11/ncu.log
