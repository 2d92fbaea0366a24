"""GPU parity: the CUDA FASTA path (through the C ABI) against the CPU oracle, bit for bit."""
import os

import numpy as np
import pytest

from tools import synth

import exb_testutil as util

pytestmark = pytest.mark.gpu


def _check_text(dev, text):
    from exon_duckdb_b200 import device as D
    from oracle import oracle as O

    buf = D.to_device(text, dev)
    try:
        want = O.parse_fasta(text)
    except O.OracleError as e:
        with pytest.raises(D.FormatError) as gi:
            D.fasta_table(buf)
        assert gi.value.pos == e.pos
        return None
    tab = D.fasta_table(buf)
    assert tab["__n_rows__"] == want.n
    for name in want.names:
        off, dat = want.column(name)
        col = tab[name]
        assert np.array_equal(col.offsets.cpu().numpy(), off), name
        assert col.data.cpu().numpy().tobytes() == dat.tobytes(), name
    assert np.array_equal(tab["description"].valid.cpu().numpy().astype(bool), want.desc_valid)
    # gc_content per record straight from the scan's prefix arrays
    s = tab["__scan__"]
    got = D.gc_from_prefix(s.seq_off, s.gc_prefix, want.n).cpu().numpy()
    ref = np.array([O.gc_content(x) for x in want.strings("sequence")], dtype=np.float32)
    assert got.tobytes() == ref.tobytes()
    assert int(s.result.gc_total) == sum(O.gc_count(x) for x in want.strings("sequence"))
    return want


@pytest.mark.parametrize("name,n", [("test.fasta", 2), ("test.mixed-desc.fasta", 2)])
def test_reference_fixtures(cuda_device, golden_dir, name, n):
    want = _check_text(cuda_device, open(os.path.join(golden_dir, name), "rb").read())
    assert want.n == n  # test_fasta_scan.test:5-8


@pytest.mark.parametrize("seed,n,kw", [
    (1, 1, {}), (2, 9, {}), (3, 400, {}), (4, 400, dict(crlf=True)), (5, 400, dict(final_eol=False)),
    (6, 3000, dict(max_len=30, wrap=7)), (7, 40, dict(min_len=20000, max_len=90000)), (8, 300, dict(max_len=0)),
    (9, 200, dict(wrap=0, max_len=5000)), (10, 500, dict(crlf=True, final_eol=False, wrap=13)),
    (11, 60, dict(min_len=16384 - 70, max_len=16384 + 70, wrap=61)), (12, 2000, dict(tricky=False, wrap=80)),
])
def test_random_records(cuda_device, seed, n, kw):
    text, _ = util.random_fasta(seed, n, **kw)
    _check_text(cuda_device, text)


def test_tile_edge_sweep(cuda_device):
    body, _ = util.random_fasta(21, 150, min_len=50, max_len=400)
    for pad in list(range(0, 70)) + [2047, 2048, 2049, 16384 - 70, 16384 - 9, 16384 - 1, 16384]:
        first = util.fasta_text([(b"pad x", b"A" * pad)], wrap=60)
        _check_text(cuda_device, first + body)


@pytest.mark.parametrize("text", [
    b"", b">a\nACGT", b">a\nACGT\n", b">a d\r\nAC\r\nGT\r\n", b">a\nAC\r", b">a", b">a\n", b">a\n>b\n>c d\n",
    b">a\n\n\nAC\n\nGT\n\n>b \nTT", b">a\tdesc with  spaces \t\nAC>GT\n>b\n>\n",
    b"ACGT\n>a\nAC\n",   # data before the first '>'
    b"\n>a\nAC\n",
    b"> a\nAC\n",        # empty name
    b">a\nAC\n>\nGG\n",  # empty name later
    b">a\r\nA\rC\r\n",   # a lone CR inside a line is data
])
def test_edge_cases(cuda_device, text):
    _check_text(cuda_device, text)


def test_header_heavy_input(cuda_device):
    _check_text(cuda_device, b"".join(b">s%d\nA\n" % i for i in range(30000)))


def test_one_long_contig_and_header_spanning_tiles(cuda_device):
    import random
    rng = random.Random(9)
    recs = [(b"chr1 " + b"x" * 40000, util.rand_seq(rng, 300000)), (b"chr2", util.rand_seq(rng, 10)),
            (b"chr3 tail", util.rand_seq(rng, 70000, b"ACGTN"))]
    _check_text(cuda_device, util.fasta_text(recs, wrap=60))
    _check_text(cuda_device, util.fasta_text(recs, wrap=0))


def test_generated_genome_matches_oracle(cuda_device):
    from exon_duckdb_b200 import _lib, device as D
    p = synth.gen_params("fasta", 64, seed=3, len_min=1000, len_max=120000)
    host = synth.gen_host(p)
    buf = synth.gen_device(p, cuda_device)
    assert buf.cpu().numpy().tobytes() == host.tobytes()
    _check_text(cuda_device, host.tobytes())
