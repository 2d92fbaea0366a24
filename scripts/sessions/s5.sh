set -x
mkdir -p gpurun_out/s5
python - <<'PY'
import sys
sys.path.insert(0,'.')
from tools import synth
synth.gen_host(synth.gen_params("illumina", 9_000_000, seed=20)).tofile('/dev/shm/big.fastq')
PY
timeout 600 ./build/rt/iobench2 /dev/shm/big.fastq > gpurun_out/s5/iobench2.txt 2>&1
cat gpurun_out/s5/iobench2.txt
timeout 1200 python -m pytest tests/test_gpu_reader2.py tests/test_duckdb_ext.py -m gpu -x -q > gpurun_out/s5/gputest.txt 2>&1
tail -30 gpurun_out/s5/gputest.txt
