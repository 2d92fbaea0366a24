set -x
mkdir -p gpurun_out/s12
for V in "X=0" "EXB_SPLIT_PREFETCH=0" "EXB_SPLIT_WIN_KB=32" "EXB_SPLIT_WIN_KB=48" "EXB_SPLIT_WIN_KB=64" "EXB_SPLIT_WIN_KB=100"; do
  echo "== $V"; env $V EXB_PATHS_SPLIT_ONLY=1 python scripts/bench_paths.py --only c2 --out /tmp/x.json 2>&1 | grep -v "^+"
done > gpurun_out/s12/split_variants.txt 2>&1
cat gpurun_out/s12/split_variants.txt
bash scripts/gpu_sanitize.sh gpurun_out/s12/sanitize > gpurun_out/s12/sanitize.txt 2>&1
tail -20 gpurun_out/s12/sanitize.txt
