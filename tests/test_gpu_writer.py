"""FASTQ / FASTA writers: the CUDA formatting kernels and the host writer object against the
oracle's record-at-a-time writers, and as the inverse of the scan.  The reference's own writer tests
(test_fastq_copy.test, test_fasta_copy.test) are commented out there; their statements are replayed through SQL in
tests/test_duckdb_ext.py."""
import gzip
import os
import random

import numpy as np
import pytest
import torch

from exb_testutil import rand_qual, rand_seq
from exon_duckdb_b200 import _lib, device as D
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _dev_col(strings, dev):
    off, data, valid = O.pack_strings(strings)
    return D.Column(torch.from_numpy(off).to(dev), D.to_device(data[: int(off[-1])], dev), torch.from_numpy(valid).to(dev))


def _records(seed, n, max_len=300, long_every=0):
    rng = random.Random(seed)
    names, descs, seqs, quals = [], [], [], []
    for i in range(n):
        L = rng.randint(0, max_len)
        if long_every and i % long_every == long_every - 1:
            L = rng.randint(70_000, 200_000)  # above the kernels' LONG_ROW: the sliced pass
        names.append(b"r%d" % i + rand_seq(rng, rng.randint(0, 20), b"abcXYZ:/_-0123456789"))
        r = rng.random()
        descs.append(None if r < 0.3 else (b"" if r < 0.4 else rand_seq(rng, rng.randint(1, 40), b"abc  :=/ GC@+>")))
        seqs.append(rand_seq(rng, L, b"ACGTN") if L < 5000 else (rand_seq(rng, 997, b"ACGTN") * (L // 997 + 1))[:L])
        quals.append(rand_qual(rng, L) if L < 5000 else (rand_qual(rng, 991) * (L // 991 + 1))[:L])
    return names, descs, seqs, quals


@pytest.mark.parametrize("seed,n,long_every", [(1, 0, 0), (2, 1, 0), (3, 257, 0), (4, 5000, 0), (5, 300, 37)])
def test_fastq_format_matches_the_oracle(cuda_device, seed, n, long_every):
    names, descs, seqs, quals = _records(seed, n, long_every=long_every)
    want = O.format_fastq(names, descs, seqs, quals)
    cols = [_dev_col(c, cuda_device) for c in (names, descs, seqs, quals)]
    img, row_off = D.fastq_format(*cols)
    assert img.cpu().numpy().tobytes() == want
    off = row_off.cpu().numpy()
    assert off[0] == 0 and off[-1] == len(want)
    if n:
        # every record starts where the previous one ended, with '@'
        starts = np.frombuffer(want, np.uint8)[off[:-1]]
        assert (starts == ord("@")).all()
    # and the image parses back to the same rows (NULL and '' descriptions both come back as NULL, as the reader defines)
    back = O.parse_fastq(want)
    assert back.strings("name") == names and back.strings("sequence") == seqs and back.strings("quality_scores") == quals
    assert back.strings("description") == [d if d else None for d in descs]


@pytest.mark.parametrize("width", [1, 7, 60, 80, 4096])
def test_fasta_format_matches_the_oracle(cuda_device, width):
    rng = random.Random(width)
    lens = [0, 1, width - 1, width, width + 1, 2 * width, 2 * width + 1, 79, 80, 81, 160, 12345] + [rng.randint(0, 700) for _ in range(200)]
    lens += [150_000, 70_001]  # sliced pass
    ids = [b"c%d" % i for i in range(len(lens))]
    descs = [None if i % 3 == 0 else (b"" if i % 3 == 1 else b"len=%d x" % L) for i, L in enumerate(lens)]
    seqs = [(rand_seq(rng, 1009, b"ACGTNacgt") * (L // 1009 + 1))[:L] for L in lens]
    want = O.format_fasta(ids, descs, seqs, line_width=width)
    cols = [_dev_col(c, cuda_device) for c in (ids, descs, seqs)]
    img, row_off = D.fasta_format(*cols, line_width=width)
    assert img.cpu().numpy().tobytes() == want
    assert int(row_off[-1]) == len(want)
    back = O.parse_fasta(want)
    assert back.strings("id") == ids and back.strings("sequence") == seqs


def test_format_is_the_inverse_of_the_scan(cuda_device):
    # device columns straight from read_fastq / read_fasta go back into a canonical file image, byte for byte
    from tools import synth

    p = synth.gen_params("illumina", 20000, seed=11)
    text = synth.gen_host(p).tobytes()
    tab = D.fastq_table(D.to_device(text, cuda_device))
    img, _ = D.fastq_format(tab["name"], tab["description"], tab["sequence"], tab["quality_scores"])
    assert img.cpu().numpy().tobytes() == text
    q = synth.gen_params("fasta", 30, seed=5, len_min=1000, len_max=90000)
    ftext = synth.gen_host(q).tobytes()
    ftab = D.fasta_table(D.to_device(ftext, cuda_device))
    ref = O.parse_fasta(ftext)
    img, _ = D.fasta_format(ftab["id"], ftab["description"], ftab["sequence"], line_width=80)
    assert img.cpu().numpy().tobytes() == O.format_fasta(ref.strings("id"), ref.strings("description"), ref.strings("sequence"), 80)
    again = D.fasta_table(D.to_device(img.cpu().numpy().tobytes(), cuda_device))
    assert again["sequence"].to_pylist() == ref.strings("sequence") and again["id"].to_pylist() == ref.strings("id")


def test_a_short_output_buffer_is_an_error_not_an_overrun(cuda_device):
    names, descs, seqs, quals = _records(9, 400)
    want = O.format_fastq(names, descs, seqs, quals)
    cols = [_dev_col(c, cuda_device) for c in (names, descs, seqs, quals)]
    with pytest.raises(_lib.ExonError) as ei:
        D.fastq_format(*cols, out_cap=len(want) - 1)
    assert ei.value.code == _lib.ERR_CAPACITY
    img, _ = D.fastq_format(*cols, out_cap=len(want))  # exactly enough
    assert img.cpu().numpy().tobytes() == want


def _host_cols(cols):
    packed = [O.pack_strings(c) for c in cols]
    return [(p[0], p[1]) for p in packed], packed[1][2]


@pytest.mark.parametrize("compression,suffix", [(None, ".fastq"), (None, ".fastq.gz"), ("gzip", ".fastq.gzip"), ("zstd", ".fastq.zstd"), (None, ".fastq.zst")])
def test_writer_object_fastq(cuda_device, tmp_path, monkeypatch, compression, suffix):
    monkeypatch.setenv("EXON_B200_WRITER_FLUSH_BYTES", "200000")  # several device round trips
    names, descs, seqs, quals = _records(21, 3000, long_every=501)
    want = O.format_fastq(names, descs, seqs, quals)
    path = str(tmp_path / ("out" + suffix))
    w = D.Writer(path, "fastq", compression)
    for lo in range(0, 3000, 700):  # ragged appends
        cols, valid = _host_cols([c[lo:lo + 700] for c in (names, descs, seqs, quals)])
        w.append(cols, valid)
    rows, nbytes = w.close()
    assert rows == 3000 and nbytes == len(want)
    raw = open(path, "rb").read()
    if suffix.endswith((".gz", ".gzip")):
        assert gzip.decompress(raw) == want
    elif suffix.endswith((".zst", ".zstd")):
        # read it back through the product's own reader (zstd input is a row of the scope table)
        assert raw[:4] == b"\x28\xb5\x2f\xfd"
        from test_gpu_reader2 import _rows_of

        p2 = path if path.endswith(".zst") else path + ".zst"
        if p2 != path:
            os.link(path, p2)
        rows = _rows_of(p2, "fastq", 4)
        assert [r[0] for r in rows] == names and [r[2] for r in rows] == seqs and [r[3] for r in rows] == quals
    else:
        assert raw == want


def test_writer_object_fasta_and_force(cuda_device, tmp_path):
    ids = [b"a", b"b", b"c"]
    descs = [b"description", None, b"x y"]
    seqs = [b"ATCG", b"ATCG" * 50, b""]
    path = str(tmp_path / "out.fasta")
    cols, valid = _host_cols([ids, descs, seqs])
    w = D.Writer(path, "fasta")
    w.append(cols, valid)
    assert w.close()[0] == 3
    assert open(path, "rb").read() == O.format_fasta(ids, descs, seqs)
    # test_fasta_copy.test:43-50: an existing file is an error unless FORCE is given
    with pytest.raises(_lib.ExonError) as ei:
        D.Writer(path, "fasta")
    assert ei.value.code == _lib.ERR_IO and "exists" in str(ei.value)
    w = D.Writer(path, "fasta", force=True, line_width=60)
    w.append(cols, valid)
    w.close()
    assert open(path, "rb").read() == O.format_fasta(ids, descs, seqs, line_width=60)
    with pytest.raises(_lib.ExonError):
        D.Writer(str(tmp_path / "x.fasta"), "fasta", compression="brotli")
    with pytest.raises(_lib.ExonError):
        D.Writer(str(tmp_path / "x.gff"), "gff")
