"""Seeded builders of FASTA/FASTQ test inputs (shared by CPU and GPU tests)."""
import random

BASES = b"ACGT"


def rand_seq(rng, n, alphabet=BASES):
    return bytes(rng.choice(alphabet) for _ in range(n))


def rand_qual(rng, n, lo=33, hi=126):
    return bytes(rng.randint(lo, hi) for _ in range(n))


def fastq_text(records, eol=b"\n", final_eol=True, plus_repeat=False):
    """records: list of (name, desc or None, seq, qual) byte tuples."""
    out = bytearray()
    for i, (name, desc, seq, qual) in enumerate(records):
        out += b"@" + name
        if desc is not None:
            out += b" " + desc
        out += eol + seq + eol + b"+"
        if plus_repeat:
            out += name
        out += eol + qual
        if i + 1 < len(records) or final_eol:
            out += eol
    return bytes(out)


def random_fastq(seed, n_records, min_len=0, max_len=300, crlf=False, final_eol=True, tricky=True):
    """Random records inside the pinned domain, with the nasty-but-legal cases switched on:
    '@' / '+' as first quality character, spaces in descriptions, empty sequences, tabs in names."""
    rng = random.Random(seed)
    recs = []
    for i in range(n_records):
        L = rng.randint(min_len, max_len)
        name = b"r%d" % i + (b"\tx" if tricky and rng.random() < 0.1 else b"") + rand_seq(rng, rng.randint(0, 12), b"abcXYZ:/_-0123456789")
        r = rng.random()
        if r < 0.3:
            desc = None
        elif r < 0.4:
            desc = b""  # "@name " : empty description
        else:
            desc = rand_seq(rng, rng.randint(1, 30), b"abc  :=/ GC@+>")
        seq = rand_seq(rng, L, b"ACGTNacgtn" if tricky else BASES)
        qual = bytearray(rand_qual(rng, L))
        if tricky and L > 0 and rng.random() < 0.3:
            qual[0] = rng.choice(b"@+>")
        recs.append((name, desc, seq, bytes(qual)))
    return fastq_text(recs, b"\r\n" if crlf else b"\n", final_eol, plus_repeat=rng.random() < 0.3), recs


def fasta_text(records, wrap=60, eol=b"\n", final_eol=True, blank_lines=False, rng=None):
    """records: list of (header line without '>', seq)."""
    out = bytearray()
    for i, (hdr, seq) in enumerate(records):
        out += b">" + hdr + eol
        lines = [seq[k:k + wrap] for k in range(0, len(seq), wrap)] if wrap else [seq]
        for j, ln in enumerate(lines):
            out += ln
            last = i + 1 == len(records) and j + 1 == len(lines)
            if not last or final_eol:
                out += eol
            if blank_lines and rng is not None and rng.random() < 0.1 and not last:
                out += eol
    return bytes(out)


def random_fasta(seed, n_records, min_len=0, max_len=500, wrap=60, crlf=False, final_eol=True, tricky=True):
    rng = random.Random(seed)
    recs = []
    for i in range(n_records):
        L = rng.randint(min_len, max_len)
        name = b"s%d" % i + rand_seq(rng, rng.randint(0, 8), b"abc.|_-")
        r = rng.random()
        if r < 0.3:
            hdr = name
        elif r < 0.4:
            hdr = name + b" "
        elif r < 0.5:
            hdr = name + b"\t  padded description \t"
        else:
            hdr = name + b" " + rand_seq(rng, rng.randint(1, 40), b"abc =:>GC  ")
        seq = bytearray(rand_seq(rng, L, b"ACGTNacgtn" if tricky else BASES))
        if tricky and L > 3 and rng.random() < 0.3:
            seq[rng.randint(1, L - 1)] = ord(">")  # '>' that is not at a line start... unless wrap puts it there
        recs.append((hdr, bytes(seq)))
    eol = b"\r\n" if crlf else b"\n"
    text = fasta_text(recs, wrap, eol, final_eol, blank_lines=tricky, rng=rng)
    return text, recs


# ---------------------------------------------------------------- host stand-ins for the shard exchange (tests only)
def host_shard_state(data, lo, hi, begin=16):
    """What a shard's scan reports about bytes [lo, hi) of `data`, computed on the host with numpy: the stand-in the CPU
    tests feed into exon_duckdb_b200.dist's composition rules (the product computes the same row on the device)."""
    import numpy as np

    from exon_duckdb_b200 import dist

    v = np.frombuffer(bytes(data[lo:hi]), dtype=np.uint8)
    nl = np.flatnonzero(v == 10)
    open_start = lo + (int(nl[-1]) + 1 if nl.size else 0)
    tail = v[open_start - lo:]
    tail_s = int(tail.astype(np.int8).astype(np.int64).sum())
    tail_g = int(((tail == ord("G")) | (tail == ord("C"))).sum())
    flags = 0
    if open_start < hi:
        b = data[open_start]
        flags = 2 if b == ord("@") else (1 if b == ord("+") else 0)
    ls = []
    if hi > lo and (lo == 0 or data[lo - 1] == 10):
        ls.append(lo)
    for p in nl[: dist.N_LS]:
        if lo + int(p) + 1 < hi and len(ls) < dist.N_LS:
            ls.append(lo + int(p) + 1)
    local_open = begin + (open_start - lo)
    return dist.state_row(lo, hi, begin, int(nl.size), local_open, tail_s, tail_g, flags, ls)


def sequential_state_at(data, pos):
    """Ground truth of the scan state just before byte `pos`: (lines before, start of the open line, byte sum, G/C, flags)."""
    import numpy as np

    v = np.frombuffer(bytes(data[:pos]), dtype=np.uint8)
    nl = np.flatnonzero(v == 10)
    start = int(nl[-1]) + 1 if nl.size else 0
    tail = v[start:]
    flags = 0
    if start < pos:
        flags = 2 if data[start] == ord("@") else (1 if data[start] == ord("+") else 0)
    return int(nl.size), start, int(tail.astype(np.int8).astype(np.int64).sum()), int(((tail == 71) | (tail == 67)).sum()), flags


def fastq_record_starts(data):
    """Offsets of the first byte of every record of a well-formed FASTQ (every 4th line start)."""
    starts, pos, line = [], 0, 0
    n = len(data)
    while pos < n:
        if line % 4 == 0:
            starts.append(pos)
        nxt = data.find(b"\n", pos)
        if nxt < 0:
            break
        pos = nxt + 1
        line += 1
    return starts


# ---- BGZF (blocked gzip) images for the device-inflate tests
import zlib


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return c.compress(data) + c.flush()


def bgzf_bytes(data, block=65280, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, eof_marker=True, sizes=None):
    """A BGZF file image of `data` (SAM specification 4.1): what bgzip writes."""
    out = []
    pos = 0
    k = 0
    while pos < len(data) or (pos == 0 and not out):
        n = sizes[k % len(sizes)] if sizes else block
        k += 1
        piece = data[pos:pos + n]
        pos += len(piece)
        z = raw_deflate(piece, level, strategy)
        assert len(z) + 26 <= 65536
        hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + (6).to_bytes(2, "little") + b"BC" + (2).to_bytes(2, "little") + (len(z) + 25).to_bytes(2, "little")
        out.append(hdr + z + zlib.crc32(piece).to_bytes(4, "little") + len(piece).to_bytes(4, "little"))
        if not data:
            break
    if eof_marker:
        out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)
