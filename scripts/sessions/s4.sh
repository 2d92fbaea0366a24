set -x
mkdir -p gpurun_out/s4
timeout 1200 python -m pytest tests/test_gpu_reader2.py tests/test_duckdb_ext.py -m gpu -x -q > gpurun_out/s4/gputest.txt 2>&1
tail -30 gpurun_out/s4/gputest.txt
EXON_B200_TRACE=1 python scripts/bench_reader.py --out gpurun_out/s4/reader.json > gpurun_out/s4/reader.txt 2>&1
grep -v "^exon_b200 reader" gpurun_out/s4/reader.txt
grep "^exon_b200 reader" gpurun_out/s4/reader.txt | head -24
