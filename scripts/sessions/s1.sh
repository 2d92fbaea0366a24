set -x
mkdir -p gpurun_out/s1
(lscpu | head -40; echo; cat /sys/devices/system/node/online; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); grep MemTotal $n/meminfo; done; nvidia-smi topo -m; nvidia-smi --query-gpu=index,pci.bus_id --format=csv; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; df -h /dev/shm /tmp; mount | grep -E "shm|tmp"; free -g; cat /sys/kernel/mm/transparent_hugepage/shmem_enabled) > gpurun_out/s1/topo.txt 2>&1
python - <<'PY'
import sys, os
sys.path.insert(0, '.')
from exon_duckdb_b200 import _lib, device as D
D.gen_host(_lib.gen_params("illumina", 6_000_000, seed=20)).tofile('/dev/shm/io.fastq')
print(os.path.getsize('/dev/shm/io.fastq'))
PY
./build/rt/iobench /dev/shm/io.fastq 0 > gpurun_out/s1/iobench.txt 2>&1
cat gpurun_out/s1/iobench.txt
EXON_B200_TRACE=1 python scripts/bench_reader.py --out gpurun_out/s1/reader.json > gpurun_out/s1/reader.txt 2>&1
tail -20 gpurun_out/s1/reader.txt
