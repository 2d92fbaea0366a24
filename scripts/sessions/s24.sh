set -x
D=gpurun_out/${1:-s24}
mkdir -p $D
timeout 1200 python -m pytest tests -m gpu -x -q > $D/gputest.txt 2>&1
tail -5 $D/gputest.txt
timeout 900 python bench.py > $D/bench.json 2> $D/bench.err
tail -3 $D/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_ref.json 2>> $D/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-c5 --no-paths > $D/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/paths_launches.csv \
    python scripts/bench_paths.py --out $D/paths_under_ncu.json > $D/ncu_paths.log 2>&1
python scripts/bench_paths.py --out $D/paths.json 2>&1 | grep -v "^+" > $D/paths.txt
python scripts/bench_reader.py --out $D/reader.json > $D/reader.txt 2>&1
python scripts/bench_duckdb.py --out $D/duckdb.json > $D/duckdb.txt 2>&1
tail -14 $D/duckdb.txt
