set -x
mkdir -p gpurun_out/s2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2/gputest.txt 2>&1
tail -15 gpurun_out/s2/gputest.txt
EXON_B200_TRACE=1 python scripts/bench_reader.py --out gpurun_out/s2/reader.json > gpurun_out/s2/reader.txt 2>&1
grep -v "^exon_b200 reader" gpurun_out/s2/reader.txt | tail -12
grep "^exon_b200 reader" gpurun_out/s2/reader.txt | sed -n '1p;4p;7p;10p;13p;16p;19p;22p'
