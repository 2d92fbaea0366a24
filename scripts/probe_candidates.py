"""One GPU: the scan + combine of the plain fused COUNT next to the scan + candidates kernel of a shard (ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from exon_duckdb_b200 import device as D, dist as XD

dev = torch.device("cuda:0")
R = int(os.environ.get("EXB_PROBE_READS", "4000000"))
buf = synth.gen_device(synth.gen_params("illumina", R, seed=20, len_min=150, len_max=150), dev)
preds = [("mean_quality", ">", 30.0)]
c = D.fastq_scan_filter(buf, preds)
torch.cuda.synchronize()
n = buf.numel()
cut = (n // 2) & ~15
shards = [XD.Shard(buf[:cut], 0, cut, 0, False), XD.Shard(buf[cut - XD.HALO:], cut, n, XD.HALO, True)]
ranges = [[s.lo, s.hi, s.begin] for s in shards]
jobs = [XD.ShardedFastqCount(s, preds, None, ranges=ranges) for s in shards]
for it in range(3):
    D.fastq_scan_filter(buf, preds, out=c)
    recs = torch.cat([j.scan_candidates().clone() for j in jobs])
    tot = jobs[-1].combine(recs)
torch.cuda.synchronize()
print(c.agg.tolist()[:5], tot.tolist()[:5])
for name, fn in (("plain", lambda: D.fastq_scan_filter(buf, preds, out=c)), ("shard0 candidates", lambda: jobs[0].scan_candidates()), ("shard1 candidates", lambda: jobs[1].scan_candidates())):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-20s %.3f ms" % (name, e0.elapsed_time(e1) / 10))
