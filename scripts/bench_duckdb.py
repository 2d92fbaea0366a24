#!/usr/bin/env python
"""The drop-in path end to end: SQL through DuckDB v0.8.1 with `LOAD exon`, file on tmpfs -> query result.

Two extensions run the same statements (tests/test_duckdb_ext.py explains them):
  PRODUCT  exon_duckdb_b200/duckdb_ext/exon.duckdb_extension  our host code over exb_reader_* (projection / filter /
           complex-filter push-down, COUNT(*) on the device, CUDA scalar functions)
  REFGLUE  oracle/_ref/exon.duckdb_extension                   the reference's UNMODIFIED C++ glue and CPU scalar
           functions over libexon_b200.so's new_reader (Arrow stream): what a user of the reference gets by
           swapping only the reader library
usage (GPU box): python scripts/bench_duckdb.py [--reads N] [--out gpurun_out/duckdb.json]
Times are DuckDB's own wall clock per statement (tools/sqlrun.cpp), second run of each statement (warm pinned pool).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SQLRUN = os.path.join(ROOT, "build", "rt", "sqlrun")
PRODUCT = os.path.join(ROOT, "exon_duckdb_b200", "duckdb_ext", "exon.duckdb_extension")
REFGLUE = os.path.join(ROOT, "oracle", "_ref", "exon.duckdb_extension")


def run(ext, statements, threads=None):
    text = "LOAD '%s';\n" % ext + "\n".join(s + ";" for s in statements) + "\n"
    cmd = [SQLRUN] + (["-threads", str(threads)] if threads else [])
    out = subprocess.run(cmd, input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1200)
    if out.returncode != 0:
        raise RuntimeError(out.stderr.decode()[-2000:])
    return [json.loads(l) for l in out.stdout.decode().splitlines()][1:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--contigs", type=int, default=2400)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--ref-reads", type=int, default=400_000, help="smaller file for the queries that run the reference's CPU scalar functions")
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--out", default="gpurun_out/duckdb.json")
    args = ap.parse_args()
    for p in (SQLRUN, PRODUCT, REFGLUE):
        if not os.path.exists(p):
            raise SystemExit("%s is missing (bash oracle/build_ref.sh; make -C exon_duckdb_b200/duckdb_ext)" % p)
    from exon_duckdb_b200 import _lib, device as D
    from tools import synth

    big = os.path.join(args.dir, "exb_duck_big.fastq")
    small = os.path.join(args.dir, "exb_duck_small.fastq")
    synth.gen_host(synth.gen_params("illumina", args.reads, seed=20)).tofile(big)
    synth.gen_host(synth.gen_params("illumina", args.ref_reads, seed=20)).tofile(small)
    fasta = os.path.join(args.dir, "exb_duck.fasta")
    synth.gen_host(synth.gen_params("fasta", args.contigs, seed=3, len_min=500000, len_max=500000, wrap=60)).tofile(fasta)
    mq = "list_avg(quality_score_string_to_list(quality_scores)) > 30"
    queries = [
        ("COUNT(*)", "SELECT COUNT(*) FROM read_fastq('%s')"),
        ("COUNT(*) WHERE mean quality > 30", "SELECT COUNT(*) FROM read_fastq('%s') WHERE " + mq),
        ("SUM(length(sequence))", "SELECT SUM(length(sequence)) FROM read_fastq('%s')"),
        ("COUNT(*) WHERE name = one record", "SELECT COUNT(*) FROM read_fastq('%s') WHERE name = 'SIM:1:FC1:1:1:1000:1000'"),
        ("AVG(gc_content(sequence))", "SELECT AVG(gc_content(sequence)) FROM read_fastq('%s')"),
        ("SUM(length(reverse_complement(sequence)))", "SELECT SUM(length(reverse_complement(sequence))) FROM read_fastq('%s')"),
        ("all 4 columns: COUNT(name), COUNT(description), SUM(length(sequence)), SUM(length(quality_scores))",
         "SELECT COUNT(name), COUNT(description), SUM(length(sequence)), SUM(length(quality_scores)) FROM read_fastq('%s')"),
        ("name, sequence WHERE mean quality > 30 -> COUNT(name), SUM(length(sequence))",
         "SELECT COUNT(name), SUM(length(sequence)) FROM read_fastq('%s') WHERE " + mq),
        ("AVG(list_avg(quality_score_string_to_list(quality_scores)))", "SELECT AVG(list_avg(quality_score_string_to_list(quality_scores))) FROM read_fastq('%s')"),
        ("SUM(len(quality_score_string_to_list(quality_scores)))", "SELECT SUM(len(quality_score_string_to_list(quality_scores))) FROM read_fastq('%s')"),
        ("C3: SELECT id, gc_content(sequence) FROM read_fasta -> SUM", "SELECT COUNT(id), SUM(gc_content(sequence)) FROM read_fasta('%s')"),
        ("C3: read_fasta all columns -> SUM(length(sequence))", "SELECT COUNT(id), SUM(length(sequence)) FROM read_fasta('%s')"),
    ]
    rows = []
    for label, ext in (("PRODUCT", PRODUCT), ("REFGLUE", REFGLUE)):
        for name, q in queries:
            # the reference's CPU scalar functions run at 0.01-0.2 GB/s per core: give them the small file
            heavy = label == "REFGLUE" and ("quality" in q or "gc_content" in q or "reverse_complement" in q)
            path = fasta if "read_fasta" in q else (small if heavy else big)
            if label == "REFGLUE" and ("len(quality" in q or "list_avg" in name and "AVG(" in name):
                continue
            size = os.path.getsize(path)
            # PRODUCT: three statements back to back (the file's page cache is copied through pinned blocks), then a pause -- an
            # interactive user's think time, in which the reader registers the page cache -- and the statement twice more
            stmts = [q % path, q % path, q % path] + ([".sleep 1500", q % path, q % path] if label == "PRODUCT" else [])
            try:
                res = run(ext, stmts, threads=args.threads or None)
            except Exception as e:
                print("%-8s %-44s FAILED %s" % (label, name, str(e)[:200]), flush=True)
                continue
            for tag, part in ((label, res[1:3]), (label + " after a pause", res[3:])):
                if not part:
                    continue
                r = min((x for x in part if x.get("ok")), key=lambda x: x["ms"], default=part[-1])
                if not r.get("ok"):
                    print("%-8s %-44s ERROR %s" % (tag, name, r.get("error", "")[:200]), flush=True)
                    continue
                ms = r["ms"]
                gbs = size / (ms * 1e-3) / 1e9
                rows.append({"extension": tag, "query": name, "file_bytes": size, "ms": ms, "GB/s": gbs, "result": r["rows"][0][0]})
                print("%-22s %-60s %9.1f ms  %7.2f GB/s  (%.2f GB file)  -> %s" % (tag, name, ms, gbs, size / 1e9, r["rows"][0][0]), flush=True)
    # ---- bgzip'ed FASTQ: inflated on the device by PRODUCT and, through new_reader, for the reference's glue;
    # EXON_B200_BGZF=0 = the streaming zlib decoder (the reference's own arrangement) on a smaller file
    from tools import paths as P
    bz = os.path.join(args.dir, "exb_duck_big.fastq.gz")
    with open(big, "rb") as f:
        text = f.read()
    with open(bz, "wb") as f:
        f.write(P.bgzf_image(text))
    sz = os.path.join(args.dir, "exb_duck_small.fastq.gz")
    with open(small, "rb") as f:
        stext = f.read()
    with open(sz, "wb") as f:
        f.write(P.bgzf_image(stext))
    for label, ext, path, tbytes, env in (("PRODUCT", PRODUCT, bz, len(text), None), ("PRODUCT, zlib stream (EXON_B200_BGZF=0)", PRODUCT, sz, len(stext), "0")):
        for name, q in queries[:2] + queries[6:7]:
            if env is not None:
                os.environ["EXON_B200_BGZF"] = env
            try:
                res = run(ext, [q % path, q % path, q % path], threads=args.threads or None)
            except Exception as e:
                print("%-8s %-44s FAILED %s" % (label, name, str(e)[:200]), flush=True)
                continue
            finally:
                os.environ.pop("EXON_B200_BGZF", None)
            r = min((x for x in res[1:] if x.get("ok")), key=lambda x: x["ms"], default=res[-1])
            if not r.get("ok"):
                print("%-8s %-44s ERROR %s" % (label, name, r.get("error", "")[:200]), flush=True)
                continue
            ms = r["ms"]
            size = os.path.getsize(path)
            rows.append({"extension": label, "query": "bgzf: " + name, "file_bytes": size, "text_bytes": tbytes, "ms": ms, "GB/s": size / ms / 1e6,
                         "text GB/s": tbytes / ms / 1e6, "result": r["rows"][0][0]})
            print("%-8s bgzf: %-54s %9.1f ms  %7.2f GB/s of file bytes = %7.2f GB/s of text (%.2f GB file)  -> %s" %
                  (label, name[:54], ms, size / ms / 1e6, tbytes / ms / 1e6, size / 1e9, r["rows"][0][0]), flush=True)
    del text, stext
    os.unlink(bz)
    os.unlink(sz)
    os.unlink(big)
    os.unlink(small)
    os.unlink(fasta)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
