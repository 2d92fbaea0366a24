set -x
D=gpurun_out/${1:-s32}
mkdir -p $D
python scripts/probe_candidates.py 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/cand_launches.csv python scripts/probe_candidates.py > $D/ncu.log 2>&1
python profiles/launch_list.py $D/cand_launches.csv | head -20
