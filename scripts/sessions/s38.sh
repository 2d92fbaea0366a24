set -x
D=gpurun_out/${1:-s38}
mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_writer.py -m gpu -x -q > $D/gputest.txt 2>&1
tail -3 $D/gputest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/paths_launches.csv \
    python scripts/bench_paths.py --only c2,c3 --out $D/paths_under_ncu.json > $D/ncu_paths.log 2>&1
python profiles/launch_list.py $D/paths_launches.csv | grep "format"
python scripts/bench_paths.py --only c2,c3 --out $D/paths.json 2>&1 | grep -v "^+" | grep "writer"
