set -x
mkdir -p gpurun_out/s10
nvidia-smi topo -m > gpurun_out/s10/topo.txt 2>&1; nproc >> gpurun_out/s10/topo.txt
timeout 900 python -m pytest tests/test_duckdb_ext.py tests/test_gpu_dist.py tests/test_gpu_fastq.py -m gpu -x -q > gpurun_out/s10/gputest.txt 2>&1
tail -5 gpurun_out/s10/gputest.txt
python - <<'PY'
import sys
sys.path.insert(0,'.')
from tools import synth
synth.gen_host(synth.gen_params("illumina", 12_000_000, seed=20)).tofile('/dev/shm/mg.fastq')
PY
cat > /tmp/mg.sql <<SQL
LOAD '$PWD/exon_duckdb_b200/duckdb_ext/exon.duckdb_extension';
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq', gpus=1);
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq', gpus=1);
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq', gpus=2);
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq', gpus=2);
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq', gpus=1) WHERE list_avg(quality_score_string_to_list(quality_scores)) > 30;
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq', gpus=2) WHERE list_avg(quality_score_string_to_list(quality_scores)) > 30;
SELECT COUNT(*) FROM read_fastq('/dev/shm/mg.fastq') WHERE list_avg(quality_score_string_to_list(quality_scores)) > 30;
SELECT COUNT(name), SUM(length(sequence)), AVG(gc_content(sequence)) FROM read_fastq('/dev/shm/mg.fastq', gpus=1);
SELECT COUNT(name), SUM(length(sequence)), AVG(gc_content(sequence)) FROM read_fastq('/dev/shm/mg.fastq', gpus=2);
SELECT COUNT(name), COUNT(description), SUM(length(sequence)), SUM(length(quality_scores)) FROM read_fastq('/dev/shm/mg.fastq', gpus=1);
SELECT COUNT(name), COUNT(description), SUM(length(sequence)), SUM(length(quality_scores)) FROM read_fastq('/dev/shm/mg.fastq', gpus=2);
SQL
./build/rt/sqlrun < /tmp/mg.sql > gpurun_out/s10/sql.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/s10/sql.txt'):
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r.get('ok'), r.get('rows'), "%.1f ms"%r.get('ms',0), r.get('error','')[:200])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s10/bench_n2.json 2> gpurun_out/s10/bench_n2.err
tail -3 gpurun_out/s10/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s10/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d.get('e2e')); print(d.get('e2e_pinned_image'))
PY
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s10/bench_n2.json').read().strip().splitlines()[-1])
print('c5',{k:v for k,v in (d.get('c5') or {}).items() if k not in ('kernels','workload')}); print(d.get('exchange'))
PY
EXB_EXCHANGE_FUSED=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-c5 > gpurun_out/s10/bench_n2_old.json 2> gpurun_out/s10/bench_n2_old.err
python - <<'PY'
import json
for f in ('gpurun_out/s10/bench_n2.json','gpurun_out/s10/bench_n2_old.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'])
PY
python bench.py --steps 10 --warmup 3 --no-e2e --no-c5 --no-paths --no-cpu > gpurun_out/s10/bench_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/s10/bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
