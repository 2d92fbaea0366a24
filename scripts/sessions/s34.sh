set -x
D=gpurun_out/${1:-s34}
mkdir -p $D
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29541 WORLD_SIZE=2 LOCAL_WORLD_SIZE=2
for R in 0 1; do
  OTHER=$((1-R))
  RANK=$OTHER LOCAL_RANK=$OTHER python bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-c5 > $D/plain_rank$OTHER.json 2> $D/plain_rank$OTHER.err &
  RANK=$R LOCAL_RANK=$R timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_rank$R.csv \
      python bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-c5 > $D/ncu_rank$R.json 2> $D/ncu_rank$R.err
  wait
  python profiles/launch_list.py $D/launches_rank$R.csv | head -14
  export MASTER_PORT=29542
done
