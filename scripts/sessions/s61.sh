#!/bin/bash
# DuckDB on a bgzip'ed FASTQ: all 4 columns, with the reader's stage trace
D=gpurun_out/s61; mkdir -p $D
python - <<'PY'
import os, sys, subprocess, json
sys.path.insert(0, '.')
from tools import synth, paths as P
text = synth.gen_host(synth.gen_params("illumina", 8_000_000, seed=20)).tobytes()
open('/dev/shm/exb_t.fastq.gz','wb').write(P.bgzf_image(text))
q = "SELECT COUNT(name), COUNT(description), SUM(length(sequence)), SUM(length(quality_scores)) FROM read_fastq('/dev/shm/exb_t.fastq.gz');"
sql = "LOAD 'exon_duckdb_b200/duckdb_ext/exon.duckdb_extension';\n" + q * 4
for env in ({}, {"EXON_B200_TRACE": "1"}, {"EXON_B200_REGISTER": "0"}):
    e = dict(os.environ); e.update(env)
    out = subprocess.run(['build/rt/sqlrun'], input=sql.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    print(env, [round(json.loads(l)['ms'],1) for l in out.stdout.decode().splitlines()[1:]])
    if env.get("EXON_B200_TRACE"): print(out.stderr.decode()[-3000:])
os.unlink('/dev/shm/exb_t.fastq.gz')
PY
