/*
 * exon_oracle.c -- CPU restatement of the reference's read_fasta / read_fastq /
 * sequence-scalar hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file.  The product (exon_duckdb_b200/) never
 * links, imports or calls it.
 *
 * Parity status
 *   - Scalar functions (gc_content, reverse_complement, complement,
 *     quality_score_string_to_list): restated from the reference's own C++
 *     (file:line cited per function) and PINNED against the reference's
 *     sqllogictest expectations (test/sql/exondb-release-with-deb-info/
 *     test_scalar_functions.test:5-46) and against vectors produced by the
 *     reference's unmodified module.cpp files compiled here (oracle/_ref,
 *     tests/golden/scalar_ref_vectors.json).
 *   - FASTA / FASTQ record parsing: the arithmetic lives in third-party crates
 *     that are NOT in /root/reference (noodles-fastq 0.8.0, noodles-fasta
 *     0.27.0, exon 0.2.6; pins rust/Cargo.lock:1274-1275,2114-2115,2127-2128;
 *     call sites rust/src/arrow_reader.rs:26-30,104-118,125-153).  The published
 *     record-at-a-time algorithm of those crates is restated below and PINNED
 *     on what the reference's tests pin (record counts, column order, the
 *     name/description split at the first space, '+' line dropped, no line
 *     terminators in strings, missing FASTA description => NULL:
 *     test_fastq_scan.test:5-8,34-41; test_fasta_scan.test:5-8,34-37;
 *     test_fasta_copy.test:74-80).  Everything else about the scan (CRLF,
 *     blank lines, TAB delimiters, truncated records) is "parity unpinned":
 *     the choice made is documented at the line that makes it.
 *
 * Plain C99, single thread, record at a time: the same shape as the
 * reference's one-scan-thread pipeline (duckdb/src/include/duckdb/function/
 * table/arrow.hpp:107-119 => max_threads = 1).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_MAX_COLS 4

typedef struct {
    int64_t n_records;
    int32_t n_cols;              /* 3 = FASTA (id, description, sequence); 4 = FASTQ */
    int64_t *off[ORC_MAX_COLS];  /* n_records + 1 Arrow-style offsets per column */
    uint8_t *data[ORC_MAX_COLS]; /* column bytes */
    int64_t cap[ORC_MAX_COLS];
    uint8_t *desc_valid;         /* one byte per record: 1 = description present */
    int64_t rec_cap;
    int64_t err_pos;             /* byte offset of the offending line, -1 if none */
    char err[128];
} orc_table;

/* ---------------------------------------------------------------- helpers */

static int orc_reserve_rows(orc_table *t, int64_t rows) {
    if (rows + 1 <= t->rec_cap) return 0;
    int64_t nc = t->rec_cap ? t->rec_cap * 2 : 1024;
    while (nc < rows + 1) nc *= 2;
    for (int c = 0; c < t->n_cols; c++) {
        int64_t *p = (int64_t *)realloc(t->off[c], (size_t)nc * sizeof(int64_t));
        if (!p) return -1;
        t->off[c] = p;
    }
    uint8_t *v = (uint8_t *)realloc(t->desc_valid, (size_t)nc);
    if (!v) return -1;
    t->desc_valid = v;
    t->rec_cap = nc;
    return 0;
}

static int orc_append(orc_table *t, int c, const uint8_t *p, int64_t len, int64_t row) {
    int64_t at = t->off[c][row];
    if (at + len > t->cap[c]) {
        int64_t nc = t->cap[c] ? t->cap[c] * 2 : 4096;
        while (nc < at + len) nc *= 2;
        uint8_t *d = (uint8_t *)realloc(t->data[c], (size_t)nc);
        if (!d) return -1;
        t->data[c] = d;
        t->cap[c] = nc;
    }
    if (len > 0) memcpy(t->data[c] + at, p, (size_t)len);
    t->off[c][row + 1] = at + len;
    return 0;
}

static void orc_init(orc_table *t, int n_cols) {
    memset(t, 0, sizeof(*t));
    t->n_cols = n_cols;
    t->err_pos = -1;
}

void orc_free_table(orc_table *t) {
    for (int c = 0; c < ORC_MAX_COLS; c++) {
        free(t->off[c]);
        free(t->data[c]);
    }
    free(t->desc_valid);
    memset(t, 0, sizeof(*t));
}

static int orc_fail(orc_table *t, int64_t pos, const char *msg) {
    t->err_pos = pos;
    snprintf(t->err, sizeof(t->err), "%s", msg);
    return -1;
}

/*
 * One line, the way noodles' read_line helpers return it (noodles-fasta 0.27
 * reader.rs read_line, noodles-fastq 0.8 reader.rs read_line): everything up to
 * '\n'; the '\n' and ONE preceding '\r' are removed.  The last line of a file
 * may lack the '\n' (then a trailing '\r' is kept, as ends_with(LF) is false).
 * Returns the offset of the next line; *len = stripped length; *had_nl says
 * whether a terminator was consumed.
 */
static int64_t orc_line(const uint8_t *b, int64_t n, int64_t pos, int64_t *len, int *had_nl) {
    const uint8_t *nl = (const uint8_t *)memchr(b + pos, '\n', (size_t)(n - pos));
    if (!nl) {
        *len = n - pos;
        *had_nl = 0;
        return n;
    }
    int64_t e = nl - b;
    int64_t l = e - pos;
    if (l > 0 && b[e - 1] == '\r') l--;
    *len = l;
    *had_nl = 1;
    return e + 1;
}

/* ------------------------------------------------------------------ FASTQ */

/*
 * noodles-fastq 0.8.0 Reader::read_record, strict 4-line records:
 *  (1) first byte must be '@' (EOF here = clean end of stream);
 *  (2) rest of the line: name = bytes before the FIRST SPACE (0x20),
 *      description = bytes after it (absent => empty)          [pinned for
 *      ' ' by test_fastq_scan.test:34-41; TAB is NOT a delimiter: unpinned];
 *  (3) next line = sequence;
 *  (4) next line must start with '+', content discarded (test2.fastq);
 *  (5) next line = quality_scores; no length check against the sequence.
 * exon 0.2.6 FASTQ array builder: empty description => NULL  [unpinned].
 * Truncated record (EOF before the quality line's first byte is reachable,
 * i.e. fewer than 4 lines) => error                          [unpinned].
 */
int orc_parse_fastq(const uint8_t *b, int64_t n, orc_table *t) {
    orc_init(t, 4);
    if (orc_reserve_rows(t, 0)) return orc_fail(t, 0, "out of memory");
    for (int c = 0; c < 4; c++) t->off[c][0] = 0;
    int64_t pos = 0, row = 0;
    while (pos < n) {
        int64_t len;
        int had_nl;
        if (b[pos] != '@') return orc_fail(t, pos, "invalid name prefix");
        if (orc_reserve_rows(t, row + 1)) return orc_fail(t, pos, "out of memory");
        /* (2) definition line */
        int64_t hs = pos + 1;
        int64_t next = orc_line(b, n, hs, &len, &had_nl);
        const uint8_t *sp = (const uint8_t *)memchr(b + hs, ' ', (size_t)len);
        int64_t name_len = sp ? (sp - (b + hs)) : len;
        int64_t desc_len = sp ? (len - name_len - 1) : 0;
        if (orc_append(t, 0, b + hs, name_len, row)) return orc_fail(t, pos, "out of memory");
        if (orc_append(t, 1, b + hs + name_len + 1, desc_len, row)) return orc_fail(t, pos, "out of memory");
        t->desc_valid[row] = desc_len > 0;
        if (!had_nl) return orc_fail(t, pos, "unexpected EOF in record");
        /* (3) sequence */
        int64_t ss = next;
        next = orc_line(b, n, ss, &len, &had_nl);
        if (orc_append(t, 2, b + ss, len, row)) return orc_fail(t, pos, "out of memory");
        if (!had_nl) return orc_fail(t, pos, "unexpected EOF in record");
        /* (4) plus line */
        int64_t ps = next;
        if (ps >= n) return orc_fail(t, pos, "unexpected EOF in record");
        if (b[ps] != '+') return orc_fail(t, ps, "invalid description prefix");
        next = orc_line(b, n, ps, &len, &had_nl);
        if (!had_nl) return orc_fail(t, pos, "unexpected EOF in record");
        /* (5) quality (the last line of the file may be unterminated) */
        int64_t qs = next;
        if (qs >= n) return orc_fail(t, pos, "unexpected EOF in record");
        next = orc_line(b, n, qs, &len, &had_nl);
        if (orc_append(t, 3, b + qs, len, row)) return orc_fail(t, pos, "out of memory");
        pos = next;
        row++;
    }
    t->n_records = row;
    return 0;
}

/* ------------------------------------------------------------------ FASTA */

/* char::is_ascii_whitespace: SPACE, TAB, LF, FF, CR (NOT VT).  Used by
 * Definition::from_str's splitn (noodles-fasta 0.27 record/definition.rs). */
static int orc_is_ascii_ws(uint8_t c) {
    return c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r';
}
/* str::trim -> char::is_whitespace restricted to one-byte code points:
 * 0x09-0x0D and SPACE.  Multi-byte Unicode spaces are outside the pinned
 * domain (generated inputs are ASCII). */
static int orc_is_trim_ws(uint8_t c) {
    return c == ' ' || (c >= 0x09 && c <= 0x0D);
}

/*
 * noodles-fasta 0.27.0 Reader::read_definition + Definition::from_str +
 * Reader::read_sequence, as driven by exon 0.2.6's FASTA batch reader:
 *  - definition = one line (LF / CRLF stripped); must start with '>';
 *    name = text up to the first ASCII whitespace, must be non-empty
 *    [pinned: test_fasta_scan.test:34-37, column `id`]; description = the
 *    remainder, str::trim()med, None when there is no whitespace at all
 *    [None => NULL documented by test_fasta_copy.test:74-80];
 *  - sequence = concatenation of the following lines (each minus LF / CRLF)
 *    until EOF or a line whose FIRST byte is '>'; blank lines add nothing.
 * Data before the first '>' is an error (read_definition: missing prefix).
 */
int orc_parse_fasta(const uint8_t *b, int64_t n, orc_table *t) {
    orc_init(t, 3);
    if (orc_reserve_rows(t, 0)) return orc_fail(t, 0, "out of memory");
    for (int c = 0; c < 3; c++) t->off[c][0] = 0;
    int64_t pos = 0, row = 0;
    while (pos < n) {
        int64_t len;
        int had_nl;
        if (b[pos] != '>') return orc_fail(t, pos, "missing prefix");
        if (orc_reserve_rows(t, row + 1)) return orc_fail(t, pos, "out of memory");
        int64_t hs = pos + 1;
        int64_t next = orc_line(b, n, hs, &len, &had_nl);
        int64_t i = 0;
        while (i < len && !orc_is_ascii_ws(b[hs + i])) i++;
        if (i == 0) return orc_fail(t, pos, "missing name");
        if (orc_append(t, 0, b + hs, i, row)) return orc_fail(t, pos, "out of memory");
        if (i < len) {
            int64_t ds = hs + i + 1, de = hs + len;
            while (ds < de && orc_is_trim_ws(b[ds])) ds++;
            while (de > ds && orc_is_trim_ws(b[de - 1])) de--;
            if (orc_append(t, 1, b + ds, de - ds, row)) return orc_fail(t, pos, "out of memory");
            t->desc_valid[row] = 1;
        } else {
            if (orc_append(t, 1, b, 0, row)) return orc_fail(t, pos, "out of memory");
            t->desc_valid[row] = 0;
        }
        t->off[2][row + 1] = t->off[2][row];
        pos = next;
        while (pos < n && b[pos] != '>') {
            int64_t ls = pos;
            pos = orc_line(b, n, ls, &len, &had_nl);
            /* append to the record's sequence: continue at its current end */
            int64_t save = t->off[2][row];
            t->off[2][row] = t->off[2][row + 1];
            int rc = orc_append(t, 2, b + ls, len, row);
            t->off[2][row] = save;
            if (rc) return orc_fail(t, ls, "out of memory");
        }
        row++;
    }
    t->n_records = row;
    return 0;
}

/* ---------------------------------------------------------- scalar functions */

/* exon/src/exon/sequence_functions/module.cpp:131-158 (per-row formula; the
 * CONSTANT_VECTOR chunk-collapse bug, SURVEY finding 4, is NOT reproduced):
 * '' -> 0.0f; else (float)#{'G','C'} / (float)size, both C ints -> float. */
float orc_gc_content(const uint8_t *s, int64_t len) {
    if (len == 0) return 0.0f;
    int64_t gc = 0;
    for (int64_t i = 0; i < len; i++) gc += (s[i] == 'G' || s[i] == 'C');
    return (float)gc / (float)len;
}

int64_t orc_gc_count(const uint8_t *s, int64_t len) {
    int64_t gc = 0;
    for (int64_t i = 0; i < len; i++) gc += (s[i] == 'G' || s[i] == 'C');
    return gc;
}

/* module.cpp:30-69: per byte, in place order, A->C T->G C->A G->T; anything
 * else throws InvalidInputException("Invalid character in sequence: <c>").
 * Returns -1 on success, else the index of the first invalid byte. */
int64_t orc_reverse_complement(const uint8_t *s, int64_t len, uint8_t *out) {
    for (int64_t i = 0; i < len; i++) {
        switch (s[i]) {
        case 'A': out[i] = 'C'; break;
        case 'T': out[i] = 'G'; break;
        case 'C': out[i] = 'A'; break;
        case 'G': out[i] = 'T'; break;
        default: return i;
        }
    }
    return -1;
}

/* module.cpp:81-121: A<->T, C<->G, same error rule. */
int64_t orc_complement(const uint8_t *s, int64_t len, uint8_t *out) {
    for (int64_t i = 0; i < len; i++) {
        switch (s[i]) {
        case 'A': out[i] = 'T'; break;
        case 'T': out[i] = 'A'; break;
        case 'C': out[i] = 'G'; break;
        case 'G': out[i] = 'C'; break;
        default: return i;
        }
    }
    return -1;
}

/* module.cpp:212-249 (transcribe): T->U, A C G unchanged, same error rule. */
int64_t orc_transcribe(const uint8_t *s, int64_t len, uint8_t *out) {
    for (int64_t i = 0; i < len; i++) {
        switch (s[i]) {
        case 'T': out[i] = 'U'; break;
        case 'A': case 'C': case 'G': out[i] = s[i]; break;
        default: return i;
        }
    }
    return -1;
}

/* module.cpp:168-203 (reverse_transcribe): U->T, A C G unchanged, same error rule. */
int64_t orc_reverse_transcribe(const uint8_t *s, int64_t len, uint8_t *out) {
    for (int64_t i = 0; i < len; i++) {
        switch (s[i]) {
        case 'U': out[i] = 'T'; break;
        case 'A': case 'C': case 'G': out[i] = s[i]; break;
        default: return i;
        }
    }
    return -1;
}

/* module.cpp:260-360 (translate_dna_to_aa): the standard codon table over upper-case ACGT.  len % 3 != 0 throws
 * "Invalid sequence length: <len>" (returns -2); a codon outside the table throws "Invalid codon: <xyz>" (returns the
 * index of its first byte); success returns -1 with len / 3 bytes in out.  The length is checked before any codon. */
int64_t orc_translate(const uint8_t *s, int64_t len, uint8_t *out) {
    /* index = 16 a + 4 b + c with A=0 C=1 G=2 T=3; rows of the reference's map re-ordered accordingly */
    static const char table[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";
    if (len % 3 != 0) return -2;
    for (int64_t i = 0; i < len; i += 3) {
        int idx = 0;
        for (int k = 0; k < 3; k++) {
            int v;
            switch (s[i + k]) {
            case 'A': v = 0; break;
            case 'C': v = 1; break;
            case 'G': v = 2; break;
            case 'T': v = 3; break;
            default: return i;
            }
            idx = idx * 4 + v;
        }
        out[i / 3] = (uint8_t)table[idx];
    }
    return -1;
}

/* exon/src/exon/fastq_functions/module.cpp:32-50: `for (auto c : string)
 * push INTEGER(c - 33)` with c a (signed) char on x86-64. */
void orc_quality_to_list(const uint8_t *s, int64_t len, int32_t *out) {
    for (int64_t i = 0; i < len; i++) out[i] = (int32_t)(signed char)s[i] - 33;
}

int64_t orc_quality_sum(const uint8_t *s, int64_t len) {
    int64_t q = 0;
    for (int64_t i = 0; i < len; i++) q += (int64_t)(signed char)s[i] - 33;
    return q;
}

/*
 * "mean quality <op> c" in reference SQL is
 *   list_avg(quality_score_string_to_list(q)) <op> c
 * DuckDB v0.8.1 list_avg over INTEGER = IntegerAverageOperationHugeint:
 * hugeint sum, then `(long double)sum / (long double)count` narrowed to
 * double (duckdb/src/function/aggregate/algebraic/avg.cpp, Finalize).  An
 * empty list gives NULL, and NULL <op> c filters the row out.
 * op: 0 '>', 1 '>=', 2 '<', 3 '<=', 4 '=', 5 '!='.
 */
double orc_mean_quality(const uint8_t *s, int64_t len) {
    long double sum = (long double)orc_quality_sum(s, len);
    return (double)(sum / (long double)len);
}

int orc_cmp_double(double v, int op, double c) {
    switch (op) {
    case 0: return v > c;
    case 1: return v >= c;
    case 2: return v < c;
    case 3: return v <= c;
    case 4: return v == c;
    default: return v != c;
    }
}

int orc_mean_quality_pass(const uint8_t *s, int64_t len, int op, double c) {
    if (len == 0) return 0;
    return orc_cmp_double(orc_mean_quality(s, len), op, c);
}

/* ------------------------------------------- whole-column helpers for timing */

/* Runs the full record-at-a-time path the C2 query takes in the reference:
 * parse every record, decode the quality string, average it, compare.
 * Returns the number of passing records, -1 on parse error. */
int64_t orc_fastq_count_mean_quality(const uint8_t *b, int64_t n, int op, double c,
                                     int64_t *n_records, int64_t *gc_total, int64_t *len_total) {
    orc_table t;
    if (orc_parse_fastq(b, n, &t)) {
        orc_free_table(&t);
        return -1;
    }
    int64_t pass = 0, gc = 0, tot = 0;
    for (int64_t r = 0; r < t.n_records; r++) {
        const uint8_t *q = t.data[3] + t.off[3][r];
        int64_t ql = t.off[3][r + 1] - t.off[3][r];
        pass += orc_mean_quality_pass(q, ql, op, c);
        const uint8_t *s = t.data[2] + t.off[2][r];
        int64_t sl = t.off[2][r + 1] - t.off[2][r];
        gc += orc_gc_count(s, sl);
        tot += sl;
    }
    if (n_records) *n_records = t.n_records;
    if (gc_total) *gc_total = gc;
    if (len_total) *len_total = tot;
    orc_free_table(&t);
    return pass;
}

/* ------------------------------------------------------------------ writers
 * The record writers the removed COPY TO path of the reference drove (declared at
 * exon/include/exon/fastq_functions/module.hpp:30; tests commented out in
 * test/sql/exondb-release-with-deb-info/test_fastq_copy.test and test_fasta_copy.test), restated from the published
 * behaviour of the pinned crates: noodles-fastq 0.8.0 Writer::write_record ('@' name [' ' description] LF sequence LF
 * '+' LF quality LF, the description only when it is not empty) and noodles-fasta 0.27.0 Writer (definition line
 * '>' name [' ' description], then the sequence in lines of line_base_count bases, 80 by default).
 * PARITY UNPINNED: the crates are not on disk and the reference has no live test of its writers; pinned only as the
 * inverse of the parsers above (tests/test_oracle_golden.py round trips).
 * One record at a time; returns the bytes written, or the bytes needed if `cap` is too small (nothing is cut). */
static int64_t orc_put(uint8_t *out, int64_t cap, int64_t pos, const uint8_t *p, int64_t n) {
    if (pos + n <= cap && n > 0) memcpy(out + pos, p, (size_t)n);
    return pos + n;
}
static int64_t orc_putc(uint8_t *out, int64_t cap, int64_t pos, uint8_t c) {
    if (pos + 1 <= cap) out[pos] = c;
    return pos + 1;
}
static int64_t orc_definition(uint8_t *out, int64_t cap, int64_t pos, uint8_t lead, const int64_t *const *off,
                              const uint8_t *const *data, const uint8_t *desc_valid, int64_t r) {
    const int64_t dl = (desc_valid && !desc_valid[r]) ? 0 : off[1][r + 1] - off[1][r];
    pos = orc_putc(out, cap, pos, lead);
    pos = orc_put(out, cap, pos, data[0] + off[0][r], off[0][r + 1] - off[0][r]);
    if (dl > 0) {
        pos = orc_putc(out, cap, pos, ' ');
        pos = orc_put(out, cap, pos, data[1] + off[1][r], dl);
    }
    return orc_putc(out, cap, pos, '\n');
}
int64_t orc_format_fastq(const int64_t *const *off, const uint8_t *const *data, const uint8_t *desc_valid, int64_t n_rows,
                         uint8_t *out, int64_t cap) {
    int64_t pos = 0;
    for (int64_t r = 0; r < n_rows; r++) {
        pos = orc_definition(out, cap, pos, '@', off, data, desc_valid, r);
        pos = orc_put(out, cap, pos, data[2] + off[2][r], off[2][r + 1] - off[2][r]);
        pos = orc_putc(out, cap, pos, '\n');
        pos = orc_putc(out, cap, pos, '+');
        pos = orc_putc(out, cap, pos, '\n');
        pos = orc_put(out, cap, pos, data[3] + off[3][r], off[3][r + 1] - off[3][r]);
        pos = orc_putc(out, cap, pos, '\n');
    }
    return pos;
}
int64_t orc_format_fasta(const int64_t *const *off, const uint8_t *const *data, const uint8_t *desc_valid, int64_t n_rows,
                         int line_width, uint8_t *out, int64_t cap) {
    int64_t pos = 0;
    for (int64_t r = 0; r < n_rows; r++) {
        pos = orc_definition(out, cap, pos, '>', off, data, desc_valid, r);
        const uint8_t *s = data[2] + off[2][r];
        const int64_t n = off[2][r + 1] - off[2][r];
        for (int64_t b = 0; b < n; b += line_width) {
            pos = orc_put(out, cap, pos, s + b, n - b < line_width ? n - b : line_width);
            pos = orc_putc(out, cap, pos, '\n');
        }
    }
    return pos;
}
