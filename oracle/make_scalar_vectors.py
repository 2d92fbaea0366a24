#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/scalar_ref_vectors.json from the REFERENCE's own scalar functions.

Runs the reference's unmodified sequence_functions/module.cpp and fastq_functions/module.cpp (compiled where they
lie by oracle/build_ref.sh into oracle/_ref/exon.duckdb_extension) inside the vendored DuckDB v0.8.1 through
build/rt/sqlrun, one single-row query per input (the reference's gc_content collapses multi-row chunks to row 0's
value, SURVEY finding 4; single-row queries pin the per-row formula).  Only runs in the container that has
/root/reference; the JSON it writes is committed.

usage: bash oracle/build_ref.sh && python oracle/make_scalar_vectors.py
"""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SQLRUN = os.path.join(ROOT, "build", "rt", "sqlrun")
EXT = os.path.join(ROOT, "oracle", "_ref", "exon.duckdb_extension")


def sql_str(s):
    return "'" + s.replace("'", "''") + "'"


def main():
    rng = random.Random(8)
    seqs = ["", "A", "G", "ATGC", "ATGCGC", "GGA", "gcGC", "GCN", "ATGCGCA", "ATCG", "GGGG", "AACG", "acgt", "ACGN", "ATCGQ",
            "NNNN", "C" * 1000, "G" * 999 + "A", "ACGT" * 5000, "N" * 20000 + "GC"]
    for _ in range(120):
        n = rng.choice([1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 150, 151, 255, 256, 1000, 4097])
        alpha = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTacgt", "GC", "AT", "ACGTRYKM-*"])
        seqs.append("".join(rng.choice(alpha) for _ in range(n)))
    # RNA alphabet and codon-shaped inputs for transcribe / reverse_transcribe / translate_dna_to_aa
    seqs += ["AUCG", "AUCU", "AUNN", "ATNN", "ATTT", "NNN", "ATGNNN", "ATGCGCA", "ATGCGCAA", "atg", "UUU",
             "AAAAATAACAAGATAATTATCATGACAACTACCACGAGAAGTAGCAGGTAATATTACTAGTTATTTTTCTTGTCATCTTCCTCGTGATGTTGCTGGCAACATCACCAGCTACTTCTCCTGCCACCTCCCCCGCGACGTCGCCGGGAAGATGACGAGGTAGTTGTCGTGGCAGCTGCCGCGGGAGGTGGCGGG"]
    for _ in range(60):
        n = rng.choice([3, 6, 9, 30, 33, 48, 150, 300, 999, 1000, 4098])
        alpha = rng.choice(["ACGT", "ACGT", "ACGT", "ACGU", "ACGTN"])
        seqs.append("".join(rng.choice(alpha) for _ in range(n)))
    quals = ["!'*5I~", "IIII5555", "!", "~", "#", "II", "@+>", "é", "Aé~", "ÿ!"]
    for _ in range(60):
        n = rng.choice([1, 2, 5, 16, 33, 100, 150, 301])
        quals.append("".join(chr(rng.randint(33, 126)) for _ in range(n)))

    stmts = ["LOAD %s" % sql_str(EXT)]
    plan = []
    for s in seqs:
        lit = sql_str(s)
        stmts.append("SELECT gc_content(%s)::DOUBLE" % lit); plan.append((s, "gc"))
        stmts.append("SELECT reverse_complement(%s)" % lit); plan.append((s, "rc"))
        stmts.append("SELECT complement(%s)" % lit); plan.append((s, "comp"))
        stmts.append("SELECT transcribe(%s)" % lit); plan.append((s, "tr"))
        stmts.append("SELECT reverse_transcribe(%s)" % lit); plan.append((s, "rtr"))
        stmts.append("SELECT translate_dna_to_aa(%s)" % lit); plan.append((s, "aa"))
    for q in quals:
        lit = sql_str(q)
        stmts.append("SELECT quality_score_string_to_list(%s)" % lit); plan.append((q, "qual"))
        stmts.append("SELECT list_avg(quality_score_string_to_list(%s))" % lit); plan.append((q, "mean_q"))
    out = subprocess.run([SQLRUN], input=(";\n".join(stmts) + ";\n").encode("utf-8"), stdout=subprocess.PIPE, check=True).stdout
    lines = [json.loads(l) for l in out.decode("utf-8").splitlines()]
    assert lines[0]["ok"], lines[0]
    lines = lines[1:]
    assert len(lines) == len(plan), (len(lines), len(plan))
    cases = {}
    order = []
    for (s, kind), res in zip(plan, lines):
        # the test reads case["seq"].encode("latin-1"): store the UTF-8 BYTES of the SQL literal as latin-1 text
        key = s.encode("utf-8").decode("latin-1")
        if key not in cases:
            cases[key] = {"seq": key}
            order.append(key)
        c = cases[key]
        if kind == "gc":
            assert res["ok"], res
            c["gc"] = float(res["rows"][0][0])  # float32 widened to double: exact
        elif kind in ("rc", "comp", "tr", "rtr"):
            if res["ok"]:
                c[kind] = res["rows"][0][0].encode("utf-8").decode("latin-1")
            else:
                assert "Invalid character in sequence" in res["error"], res
                c[kind] = None
        elif kind == "aa":
            if res["ok"]:
                c["aa"] = res["rows"][0][0]
            else:  # keep the reference's message: which of the two checks fired, and on what
                msg = res["error"].split("Invalid Input Error: ", 1)[-1]
                assert msg.startswith("Invalid sequence length: ") or msg.startswith("Invalid codon: "), res
                c["aa_error"] = msg.encode("utf-8").decode("latin-1")
        elif kind == "qual":
            assert res["ok"], res
            c["qual"] = json.loads(res["rows"][0][0])
        elif kind == "mean_q":
            assert res["ok"], res
            c["mean_q"] = float(res["rows"][0][0])
    doc = {"source": "reference exon/src/exon/{sequence_functions,fastq_functions}/module.cpp inside vendored DuckDB v0.8.1 "
                     "(oracle/build_ref.sh, oracle/make_scalar_vectors.py)",
           "cases": [cases[k] for k in order]}
    path = os.path.join(ROOT, "tests", "golden", "scalar_ref_vectors.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote %s: %d cases" % (path, len(doc["cases"])))


if __name__ == "__main__":
    sys.exit(main())
