#!/bin/bash
# registered page cache: parity test, reader C ABI on both I/O paths, bench e2e legs, DuckDB
D=gpurun_out/s47; mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_reader2.py tests/test_gpu_scalar_reader.py tests/test_duckdb_ext.py -m gpu -x -q > $D/pytest.txt 2>&1; echo "pytest exit $?"; tail -5 $D/pytest.txt
python scripts/bench_reader.py --out $D/reader.json > $D/reader.txt 2>&1; cat $D/reader.txt
timeout 600 python bench.py --no-paths --no-c5 --no-cpu > $D/bench.json 2> $D/bench.err; echo "bench exit $?"; tail -3 $D/bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s47/bench.json').read().strip().splitlines()[-1])
for k in ('e2e','e2e_first_scan','e2e_pinned_image'):
    print(k, {a:b for a,b in l.get(k,{}).items() if a in ('value','ms_per_step','h2d_peak_gbs','frac_of_h2d_peak','io_path')})
PY
python scripts/bench_duckdb.py --out $D/duckdb.json > $D/duckdb.txt 2>&1; tail -25 $D/duckdb.txt
