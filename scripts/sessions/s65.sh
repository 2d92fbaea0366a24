#!/bin/bash
# last check of the final tree: GPU suite, smoke, the bench line
D=gpurun_out/s65; mkdir -p $D
timeout 1200 python -m pytest tests -m gpu -x -q > $D/gputest.txt 2>&1; echo "pytest exit $?" | tee -a $D/gputest.txt; tail -3 $D/gputest.txt
python __graft_entry__.py smoke > $D/smoke.txt 2>&1; tail -1 $D/smoke.txt
timeout 900 python bench.py > $D/bench.json 2> $D/bench.err; echo "bench exit $?"; tail -2 $D/bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s65/bench.json').read().strip().splitlines()[-1])
print(l['value'], l['roofline']['frac'], l['e2e']['value'], l['e2e'].get('io_path'), l['e2e_first_scan']['value'], l['c5']['ms_per_step'])
print([ (r['path'][:30], round(r['ms_median'],2)) for r in l['paths'] if 'BGZF' in r['path']])
PY
