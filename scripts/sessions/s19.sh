set -x
D=gpurun_out/${1:-s19}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_fullsize.py tests/test_gpu_reader2.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -15 $D/gputest.txt
python bench.py --steps 20 --warmup 5 --no-e2e --no-c5 --no-paths --no-cpu > $D/bench_n1.json 2>$D/bench.err
python -c "
import json; d=json.loads(open('$D/bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d['roofline']['frac'])"
EXB_BENCH_READS=4000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastq_tile_kernel -s 3 -c 1 \
    -f -o $D/fastq_scan python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-c5 --no-paths > $D/ncu_full.log 2>&1
echo "ncu full exit $?"
python scripts/bench_paths.py --only c2,c4 --out $D/paths.json 2>&1 | grep -v "^+" | tee $D/paths.txt
