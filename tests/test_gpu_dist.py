"""Byte-range shards on the device (exon_duckdb_b200/dist.py): G shards of one file driven through the real kernels on
one GPU (LocalGroup), compared with the single-shot scan and the oracle.  The NCCL path proper runs in bench.py --gpus N."""
import random

import numpy as np
import pytest

from tools import synth
import torch

import exb_testutil as util
from exon_duckdb_b200 import _lib, device as D, dist
from oracle import oracle as O

pytestmark = pytest.mark.gpu

PREDS = [("mean_quality", ">", 30.0)]


def _sharded_count(data, cuts, dev, preds=PREDS):
    bounds = [0] + list(cuts) + [len(data)]
    G = len(bounds) - 1
    shards = []
    for k in range(G):
        lo, hi = bounds[k], bounds[k + 1]
        begin = 0 if k == 0 else dist.HALO
        halo = bytes(max(0, begin - lo)) + bytes(data[max(0, lo - begin):lo]) if begin else b""
        shards.append(dist.Shard(D.to_device(halo + bytes(data[lo:hi]), dev), lo, hi, begin, k == G - 1))
    ranges = [[s.lo, s.hi, s.begin] for s in shards]
    jobs = [dist.ShardedFastqCount(s, preds, None, ranges=ranges) for s in shards]
    blocks = torch.cat([j.scan().clone() for j in jobs])  # the all-gather
    total = torch.zeros(8, dtype=torch.int64, device=dev)
    for k, j in enumerate(jobs):
        total += j.resolve(blocks, k)  # the all-reduce
        if k:  # the device composition equals its host mirror
            want = dist.compose_prev(j.states_from(blocks.view(G, -1).cpu().numpy()), k, j.shard.begin)
            got = j.true_prev.view(torch.int64).cpu().tolist()
            assert (got[0], got[1], got[7], got[8], (got[3] >> 32) & 3) == (want.total_lines, want.open_line_start, want.tail_s, want.tail_g, want.pad)
    # the single-exchange flavour must give the same aggregates: K2 for all four phases before the "all-gather", one combine
    jobs2 = [dist.ShardedFastqCount(s, preds, None, ranges=ranges) for s in shards]  # (own workspaces: `jobs` keep their results)
    recs = torch.cat([j.scan_candidates().clone() for j in jobs2])
    fused = jobs2[-1].combine(recs).clone()
    assert bool(fused[7]) == bool(total[7]) and int(fused[6]) == int(total[6]), (fused.tolist(), total.tolist(), cuts)
    if not bool(total[7]):  # (a malformed file is an error either way; the partial sums are unspecified)
        assert fused[[0, 3, 4]].tolist() == total[[0, 3, 4]].tolist(), (fused.tolist(), total.tolist(), cuts)
    assert jobs2[0].combine(recs).tolist() == fused.tolist()
    return total, jobs


def _expect(data, op=">", thr=30.0):
    n_pass, n_rec, _, _ = O.fastq_count_mean_quality(data, op, thr)
    ref = O.parse_fastq(data)
    quals = [q for q in ref.strings("quality_scores") if O.mean_quality_pass(q, op, thr)]
    return n_pass, sum(O.quality_sum(q) for q in quals), sum(len(q) for q in quals), n_rec


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, {"crlf": True}), (3, {"final_eol": False}), (4, {"max_len": 3}), (5, {"max_len": 9000, "min_len": 3000})])
def test_sharded_count_equals_oracle(cuda_device, seed, kw):
    n = 60 if kw.get("max_len", 0) < 1000 else 12
    data, _ = util.random_fastq(seed, n, **kw)
    want = _expect(data)
    rng = random.Random(seed)
    cut_sets = [[c] for c in range(0, len(data) + 1, max(1, len(data) // 97))]
    cut_sets += [sorted(rng.randint(0, len(data)) for _ in range(rng.randint(2, 6))) for _ in range(25)]
    for cuts in cut_sets:
        total, _ = _sharded_count(data, cuts, cuda_device)
        t = dist.check_count(total)
        assert (t[0], t[3], t[4]) == want[:3], (cuts, t, want)


def test_sharded_count_generated_illumina(cuda_device):
    p = synth.gen_params("illumina", 30000, seed=20)
    data = synth.gen_host(p).tobytes()
    want = _expect(data)
    for G in (2, 3, 8):
        cuts = [dist.byte_range(len(data), k, G)[0] for k in range(1, G)]
        total, jobs = _sharded_count(data, cuts, cuda_device)
        t = dist.check_count(total)
        assert (t[0], t[3], t[4]) == want[:3]
        assert jobs[-1].c.result.total_lines == 4 * want[3]


def test_sharded_count_reports_malformed_and_truncated(cuda_device):
    data, _ = util.random_fastq(9, 30, tricky=False)
    bad = data.replace(b"\n+", b"\n-", 1)
    with pytest.raises(D.FormatError):
        dist.check_count(_sharded_count(bad, [len(bad) // 3, 2 * len(bad) // 3], cuda_device)[0])
    cut = data[: data.rfind(b"\n+")]
    with pytest.raises(D.FormatError):
        dist.check_count(_sharded_count(cut, [len(cut) // 2], cuda_device)[0])


def test_shard_state_and_record_bounds(cuda_device):
    for seed, kw in [(1, {}), (2, {"crlf": True}), (6, {"max_len": 5})]:
        data, _ = util.random_fastq(seed, 40, **kw)
        rec = util.fastq_record_starts(data)
        rng = random.Random(seed)
        for _ in range(12):
            G = rng.randint(2, 5)
            cuts = sorted(rng.randint(0, len(data)) for _ in range(G - 1))
            bounds = [0] + cuts + [len(data)]
            states = []
            for k in range(G):
                lo, hi = bounds[k], bounds[k + 1]
                begin = 0 if k == 0 else dist.HALO
                halo = bytes(max(0, begin - lo)) + bytes(data[max(0, lo - begin):lo]) if begin else b""
                sh = dist.Shard(D.to_device(halo + bytes(data[lo:hi]), cuda_device), lo, hi, begin, k == G - 1)
                row = dist.fastq_shard_state(sh)
                want = util.host_shard_state(data, lo, hi, begin)
                want[dist.TAIL_S] = want[dist.TAIL_G] = row[dist.TAIL_S] = row[dist.TAIL_G] = 0  # not computed by the line-only scan
                if want[dist.OPEN_START] >= hi:
                    want[dist.OPEN_FLAGS] = row[dist.OPEN_FLAGS] = 0  # the open line has no byte in the shard
                assert row == want, (cuts, k)
                states.append(row)
            got = dist.fastq_record_bounds(states, len(data))
            assert got == [next((r for r in rec if r >= bounds[k]), len(data)) for k in range(G)] + [len(data)]
            # every [S_k, S_{k+1}) is a complete file image: the shards' tables concatenate to the whole table
            names = []
            for k in range(G):
                part = data[got[k]:got[k + 1]]
                if part:
                    names += D.fastq_table(D.to_device(part, cuda_device), columns=["name"])["name"].to_pylist()
            assert names == O.parse_fastq(data).strings("name")


def test_fasta_first_header_and_bounds(cuda_device):
    data, recs = util.random_fasta(4, 30, max_len=400)
    hdrs = [i for i in range(len(data)) if data[i] == ord(">") and (i == 0 or data[i - 1] == 10)]
    rng = random.Random(4)
    for _ in range(20):
        G = rng.randint(2, 5)
        cuts = sorted(rng.randint(0, len(data)) for _ in range(G - 1))
        bounds = [0] + cuts + [len(data)]
        first = []
        for k in range(G):
            lo, hi = bounds[k], bounds[k + 1]
            begin = 0 if k == 0 else dist.HALO
            halo = bytes(max(0, begin - lo)) + bytes(data[max(0, lo - begin):lo]) if begin else b""
            sh = dist.Shard(D.to_device(halo + bytes(data[lo:hi]), cuda_device), lo, hi, begin, k == G - 1)
            got = dist.fasta_first_header(sh, window=64)
            assert got == next((h for h in hdrs if lo <= h < hi), -1), (cuts, k)
            first.append(got)
        b = dist.fasta_record_bounds(first, len(data))
        ids = []
        for k in range(G):
            part = data[b[k]:b[k + 1]]
            if part:
                ids += D.fasta_table(D.to_device(part, cuda_device), columns=["id"])["id"].to_pylist()
        assert ids == O.parse_fasta(data).strings("id")


# ------------------------------------------------------------------ C5: COUNT + SUM(#GC) + SUM(len) + AVG(gc_content) over byte-range shards
def _sharded_totals(data, cuts, dev, fused=True):
    bounds = [0] + list(cuts) + [len(data)]
    G = len(bounds) - 1
    shards = []
    for k in range(G):
        lo, hi = bounds[k], bounds[k + 1]
        begin = 0 if k == 0 else dist.HALO
        halo = bytes(max(0, begin - lo)) + bytes(data[max(0, lo - begin):lo]) if begin else b""
        shards.append(dist.Shard(D.to_device(halo + bytes(data[lo:hi]), dev), lo, hi, begin, k == G - 1))
    ranges = [[s.lo, s.hi, s.begin] for s in shards]
    jobs = [dist.ShardedFastqTotals(s, None, ranges=ranges, fused=fused) for s in shards]
    blocks = torch.cat([j.scan().clone() for j in jobs])  # the all-gather
    total = torch.zeros(8, dtype=torch.int64, device=dev)
    for k, j in enumerate(jobs):
        j.resolve_local(blocks, k)
        j.finish_local()
        blk = dist._result_block(j.ws)
        t = j.agg.clone()
        t[7] = (blk[2] != 0).to(t.dtype)
        t[6] = (blk[0] & 3) if j.shard.is_last else 0
        total += t  # the all-reduce
    return total


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, {"crlf": True}), (3, {"final_eol": False}), (4, {"max_len": 3}), (5, {"max_len": 9000, "min_len": 3000})])
def test_sharded_totals_equal_oracle(cuda_device, seed, kw):
    n = 60 if kw.get("max_len", 0) < 1000 else 12
    data, _ = util.random_fastq(seed, n, **kw)
    seqs = O.parse_fastq(data).strings("sequence")
    want = (len(seqs), sum(len(s) for s in seqs), sum(s.count(b"G") + s.count(b"C") for s in seqs),
            sum(int(round(float(np.float64(O.gc_content(s)) * 4294967296.0))) for s in seqs if s))
    rng = random.Random(seed)
    cut_sets = [[c] for c in range(0, len(data) + 1, max(1, len(data) // 61))]
    cut_sets += [sorted(rng.randint(0, len(data)) for _ in range(rng.randint(2, 7))) for _ in range(25)]
    for cuts in cut_sets:
        # the fused TOTALS flavour (aggregates of all four phase hypotheses in the byte pass) and the general scan +
        # exb_fastq_seq_totals over per-record arrays: both exact
        for fused in (True, False):
            t = dist.check_count(_sharded_totals(data, cuts, cuda_device, fused))
            assert (t[0], t[1], t[2], t[5]) == want, (fused, cuts, t, want)
    # one shot over the whole file
    agg, res = D.fastq_scan_totals(D.to_device(data, cuda_device))
    a = agg.cpu().tolist()
    assert res.err_pos == _lib.NO_POS and res.total_lines == 4 * want[0]
    assert (a[1], a[2], a[3], a[5]) == (want[1], want[2], want[0], want[3])
