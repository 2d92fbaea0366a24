"""Host side of the multi-GPU path (exon_duckdb_b200/dist.py) on CPU: the composition rules behind the boundary-resync
step, swept over every shard boundary of small files (the "fake cluster" of SURVEY 4), and the same rules driven through
torch.distributed with the gloo backend at world_size 2."""
import os
import random
import socket

import numpy as np
import pytest

import exb_testutil as util
from exon_duckdb_b200 import dist


def _texts():
    out = []
    for seed, kw in [(1, {}), (2, {"crlf": True}), (3, {"final_eol": False}), (4, {"min_len": 0, "max_len": 3}), (5, {"max_len": 40})]:
        out.append(util.random_fastq(seed, 12, **kw)[0])
    return out


def _check_partition(data, cuts):
    bounds = [0] + list(cuts) + [len(data)]
    G = len(bounds) - 1
    states = [util.host_shard_state(data, bounds[k], bounds[k + 1], 0 if k == 0 else dist.HALO) for k in range(G)]
    assert dist.compose_prev(states, 0) is None
    for k in range(1, G):
        lines, start, s, g, flags = util.sequential_state_at(data, bounds[k])
        prev = dist.compose_prev(states, k, dist.HALO)
        assert prev.total_lines == lines
        assert prev.open_line_start == dist.HALO + (start - bounds[k])
        assert (prev.tail_s, prev.tail_g) == (s, g)
        if start < bounds[k]:
            assert prev.pad == flags
    rec = util.fastq_record_starts(data)
    want = [next((r for r in rec if r >= bounds[k]), len(data)) for k in range(G)] + [len(data)]
    # a record start that lies beyond the shard's own end belongs to a later shard: this one is empty
    got = dist.fastq_record_bounds(states, len(data))
    assert got == want, (cuts, got, want)
    assert got[0] == 0 or not rec


def test_byte_range_partitions_the_file():
    for n in (0, 1, 7, 100, 12345):
        for world in (1, 2, 3, 8):
            r = [dist.byte_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))


@pytest.mark.parametrize("idx", range(5))
def test_two_shards_every_boundary(idx):
    data = _texts()[idx]
    for c in range(0, len(data) + 1):
        _check_partition(data, [c])


def test_many_shards_random_boundaries_including_empty_and_newline_free_shards():
    rng = random.Random(7)
    for data in _texts():
        for _ in range(150):
            G = rng.randint(3, 7)
            cuts = sorted(rng.randint(0, len(data)) for _ in range(G - 1))
            _check_partition(data, cuts)
    # shards that fall inside one long line
    data = b"@r0\n" + b"A" * 500 + b"\n+\n" + b"@" * 500 + b"\n@r1\nAC\n+\n+I\n"
    for cuts in ([10, 100, 200, 300], [4, 505], [505, 508, 600, 1009], [1, 2, 3]):
        _check_partition(data, cuts)


def test_fasta_bounds():
    assert dist.fasta_record_bounds([0, -1, 40, -1], 100) == [0, 40, 40, 100, 100]
    assert dist.fasta_record_bounds([-1, -1], 10) == [10, 10, 10]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, data, q):
    import torch
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dist.TorchGroup()
        lo, hi = dist.byte_range(len(data), rank, world)
        row = util.host_shard_state(data, lo, hi, 0 if rank == 0 else dist.HALO)
        states = g.all_gather_rows(row)
        bounds = dist.fastq_record_bounds(states, len(data))
        prev = dist.compose_prev(states, rank, dist.HALO)
        # each rank "counts" the records it owns; the all-reduce must give the file's total on every rank
        rec = util.fastq_record_starts(data)
        mine = sum(1 for r in rec if bounds[rank] <= r < bounds[rank + 1])
        agg = torch.zeros(8, dtype=torch.int64)
        agg[0] = mine
        g.all_reduce_sum(agg)
        q.put((rank, bounds, None if prev is None else (int(prev.total_lines), int(prev.open_line_start), int(prev.tail_s)), int(agg[0]), len(rec)))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_exchange(world):
    import torch.multiprocessing as mp

    data = util.random_fastq(11, 40, max_len=80)[0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, data, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    bounds = res[0][1]
    rec = util.fastq_record_starts(data)
    for rank, b, prev, total, n_rec in res:
        assert b == bounds and total == n_rec == len(rec)
        lo, _ = dist.byte_range(len(data), rank, world)
        if rank == 0:
            assert prev is None
        else:
            lines, start, s, _, _ = util.sequential_state_at(data, lo)
            assert prev == (lines, dist.HALO + start - lo, s)
    assert bounds[0] == 0 and bounds[-1] == len(data) and all(x in rec or x == len(data) for x in bounds)
