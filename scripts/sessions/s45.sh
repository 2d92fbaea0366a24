set -x
D=gpurun_out/${1:-s45}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_fullsize.py tests/test_gpu_fastq.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -12 $D/gputest.txt
EXB_BENCH_C5_GB=100 python bench.py --steps 5 --warmup 3 --no-e2e --no-paths --no-cpu > $D/bench.json 2> $D/bench.err
python -c "
import json; d=json.loads(open('$D/bench.json').read().strip().splitlines()[-1]); print({k:v for k,v in d['c5'].items() if k not in ('kernels','workload')})"
EXB_TOTALS_FUSED=0 python bench.py --steps 5 --warmup 3 --no-e2e --no-paths --no-cpu > $D/bench_general.json 2>> $D/bench.err
python -c "
import json; d=json.loads(open('$D/bench_general.json').read().strip().splitlines()[-1]); print({k:v for k,v in d['c5'].items() if k not in ('kernels','workload')})"
