set -x
D=gpurun_out/${1:-s42}
mkdir -p $D
timeout 1200 python -m pytest tests/test_gpu_reader2.py tests/test_gpu_scalar_reader.py tests/test_duckdb_ext.py tests/test_gpu_writer.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -5 $D/gputest.txt
python scripts/bench_reader.py --out $D/reader.json > $D/reader.txt 2>&1; tail -12 $D/reader.txt
EXON_B200_NO_BORROW=1 python scripts/bench_reader.py --out $D/reader_noborrow.json > $D/reader_noborrow.txt 2>&1; tail -12 $D/reader_noborrow.txt
python scripts/bench_duckdb.py --out $D/duckdb.json > $D/duckdb.txt 2>&1
grep PRODUCT $D/duckdb.txt
