#!/bin/bash
# BGZF writer sink + 200 ms idle gate: writer / DuckDB copy tests, registered-page-cache test, COPY TO gzip timing
D=gpurun_out/s62; mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_writer.py tests/test_duckdb_ext.py tests/test_gpu_reader2.py tests/test_inflate.py -m gpu -x -q > $D/pytest.txt 2>&1; echo "pytest exit $?"; tail -3 $D/pytest.txt
python - <<'PY' 2>&1 | tee gpurun_out/s62/copy_gzip.txt
import os, sys, subprocess, json, gzip
sys.path.insert(0, '.')
from tools import synth
src = '/dev/shm/exb_w.fastq'
synth.gen_host(synth.gen_params("illumina", 2_000_000, seed=20)).tofile(src)
size = os.path.getsize(src)
for mode in ("bgzf", "zlib"):
    dst = '/dev/shm/exb_w_%s.fastq.gz' % mode
    sql = "LOAD 'exon_duckdb_b200/duckdb_ext/exon.duckdb_extension';\nCOPY (FROM read_fastq('%s')) TO '%s' (FORMAT 'fastq', FORCE true);\nCOPY (FROM read_fastq('%s')) TO '%s' (FORMAT 'fastq', FORCE true);\nSELECT COUNT(*), SUM(length(sequence)) FROM read_fastq('%s');\n" % (src, dst, src, dst, dst)
    e = dict(os.environ)
    if mode == "zlib": e["EXON_B200_GZIP_WRITER"] = "zlib"
    out = subprocess.run(['build/rt/sqlrun'], input=sql.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    r = [json.loads(l) for l in out.stdout.decode().splitlines()[1:]]
    ms = r[1]['ms']
    print("COPY ... TO (FORMAT 'fastq') *.gz, %s sink: %.0f ms for %.2f GB of text = %.2f GB/s -> %.2f GB file; read back: %s in %.0f ms" %
          (mode, ms, size / 1e9, size / ms / 1e6, os.path.getsize(dst) / 1e9, r[2].get('rows'), r[2].get('ms', 0)))
    os.unlink(dst)
os.unlink(src)
PY
