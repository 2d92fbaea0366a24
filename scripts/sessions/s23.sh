set -x
D=gpurun_out/${1:-s23}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_writer.py tests/test_duckdb_ext.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -30 $D/gputest.txt
