/*
 * exon_b200.h -- C ABI of the B200-native FASTA/FASTQ scan engine.
 *
 * Plain C: pointers, sizes, PODs.  No C++ or torch types cross this boundary.
 * Every function returns 0 on success or a negative exb_status; the message for
 * the calling thread's last failure is exb_last_error().
 *
 * Two layers:
 *
 *  (1) Reference-FFI layer -- the two symbols the reference's C++ glue binds for
 *      the read_fasta / read_fastq path, same names and signatures, so the
 *      unmodified glue links against this library instead of the Rust staticlib:
 *        new_reader        <- exon/include/rust.hpp:41-46, rust/src/arrow_reader.rs:38-166
 *                             callers: arrow_table_function/module.cpp:95-103 (bind),
 *                             :235-244 (init_global)
 *        replacement_scan  <- exon/include/rust.hpp:48, rust/src/arrow_reader.rs:173-197
 *                             caller: arrow_table_function/module.cpp:324
 *      The Arrow C data / stream structs are the public Arrow ABI.
 *
 *  (2) Device layer (exb_*) -- the kernels behind it, callable on buffers that
 *      are already resident in HBM (bench.py `value`, parity tests, and hosts
 *      that keep data on the GPU).  All d_* pointers are device pointers; `stream`
 *      is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *      asynchronous on `stream` unless stated otherwise.
 */
#ifndef EXON_B200_H
#define EXON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EXB_API __attribute__((visibility("default")))
#else
#define EXB_API
#endif

typedef enum exb_status {
    EXB_OK = 0,
    EXB_ERR_CUDA = -1,      /* a CUDA call failed (no device, launch error, ...) */
    EXB_ERR_ARG = -2,       /* bad argument */
    EXB_ERR_FORMAT = -3,    /* malformed FASTA/FASTQ input (see err_pos) */
    EXB_ERR_CAPACITY = -4,  /* an output buffer was too small */
    EXB_ERR_IO = -5,        /* file could not be opened / read */
    EXB_ERR_INVALID_CHAR = -6 /* reverse_complement/complement met a byte outside ACGT */
} exb_status;

EXB_API const char *exb_last_error(void);
EXB_API const char *exb_version(void);
/* 1 if a CUDA device is usable from this process, else 0 (never fails). */
EXB_API int exb_device_available(void);

/* ---------------------------------------------------------------------------
 * (1) Reference-FFI layer
 * ------------------------------------------------------------------------- */

#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
    const char *format;
    const char *name;
    const char *metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema **children;
    struct ArrowSchema *dictionary;
    void (*release)(struct ArrowSchema *);
    void *private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void **buffers;
    struct ArrowArray **children;
    struct ArrowArray *dictionary;
    void (*release)(struct ArrowArray *);
    void *private_data;
};
#endif
#ifndef ARROW_C_STREAM_INTERFACE
#define ARROW_C_STREAM_INTERFACE
struct ArrowArrayStream {
    int (*get_schema)(struct ArrowArrayStream *, struct ArrowSchema *out);
    int (*get_next)(struct ArrowArrayStream *, struct ArrowArray *out);
    const char *(*get_last_error)(struct ArrowArrayStream *);
    void (*release)(struct ArrowArrayStream *);
    void *private_data;
};
#endif

/* rust.hpp:7-9.  error == NULL on success; otherwise a NUL-terminated message
 * the caller may release with exb_free_string (the reference leaks it). */
typedef struct ReaderResult {
    const char *error;
} ReaderResult;
/* rust.hpp:11-13.  "FASTA" / "FASTQ" or NULL. */
typedef struct ReplacementScanResult {
    const char *file_type;
} ReplacementScanResult;

/*
 * Open `uri` (a local file or a directory of files) as an Arrow stream of
 * record batches of at most `batch_size` rows (the reference passes 2048).
 *   compression: NULL = infer from the last '.'-suffix ("gz" / "zst"), else
 *                "gzip" | "gz", "zstd" | "zst", "bzip2" | "bz2", "xz" (datafusion 28
 *                FileCompressionType::from_str, case-insensitive); anything else =
 *                uncompressed (arrow_reader.rs:60-91).  gzip through zlib; zstd, bzip2
 *                and xz through the system's shared libraries, bound at run time.
 *   file_format: "fasta" or "fastq" (case-insensitive).
 *   filters:     NULL / "" or the predicate text produced by the reference's
 *                FilterToString (module.cpp:158-214): `col<op>'const'` terms
 *                joined by AND / OR, `col IS [NOT] NULL`.  This build also
 *                accepts the function predicates it can push down:
 *                  mean_quality(quality_scores) <op> number
 *                  gc_content(sequence) <op> number
 *                  length(sequence) <op> integer
 * Schema: FASTA (id, description, sequence); FASTQ (name, description,
 * sequence, quality_scores); all utf8, description nullable.
 */
EXB_API ReaderResult new_reader(struct ArrowArrayStream *stream_ptr, const char *uri, uintptr_t batch_size,
                                const char *compression, const char *file_format, const char *filters);
EXB_API ReplacementScanResult replacement_scan(const char *uri);
EXB_API void exb_free_string(const char *s);

/*
 * Native reader: the same engine as new_reader without the Arrow detour, for hosts that fill their own vectors
 * (exon_duckdb_b200/duckdb_ext/exon_extension.cpp: bind / init_global / scan callbacks).  Replaces the
 * new_reader + ArrowScanParallelStateNext + ArrowToDuckDB chain of the reference
 * (arrow_table_function/module.cpp:95-103,235-252,257-294; duckdb/src/function/table/arrow.cpp:248-268).
 *   column_mask: bit c set = column c is wanted (projection push-down); others come back as NULL views.
 *   A batch borrows the reader's host buffers: offsets[i] .. offsets[i+1] index `data` for row i
 *   (offsets are NOT rebased to 0).  The buffers stay alive until exb_batch_release, even after
 *   exb_reader_next / exb_reader_close -- a host may hand out pointers into them (DuckDB string_t).
 */
typedef struct exb_reader exb_reader;

/* Computed output columns: a scalar function of the reference applied to a column of the scan, evaluated on the
 * device while the chunk is in HBM, so that `SELECT gc_content(sequence) FROM read_fasta(..)` never ships the
 * sequence bytes to the host (north_star kernel (3); the DuckDB extension's OptimizerExtension rewrites such
 * projections into these).  Replaces the per-row loops of sequence_functions/module.cpp:30-253 and
 * fastq_functions/module.cpp:32-50 for inputs that come straight from the scan. */
#define EXB_C_GC_CONTENT 1   /* gc_content(sequence)                       -> FLOAT   (module.cpp:131-158)           */
#define EXB_C_SEQ_MAP 2      /* reverse_complement | complement | transcribe | reverse_transcribe (sequence)
                                                                           -> VARCHAR, arg = EXB_MAP_*               */
#define EXB_C_QUALITY_LIST 3 /* quality_score_string_to_list(quality_scores) -> INTEGER[] (fastq_functions:32-50)   */
#define EXB_C_MEAN_QUALITY 4 /* list_avg(quality_score_string_to_list(quality_scores)) -> DOUBLE, NULL for ''      */
#define EXB_C_SEQ_LENGTH 5   /* length(sequence)       -> BIGINT (bytes; equals DuckDB's length() for ASCII only,
                                so the extension uses it for strlen(), not length())                                */
#define EXB_C_QUAL_LENGTH 6  /* strlen(quality_scores) -> BIGINT                                                    */
typedef struct exb_computed {
    int32_t kind; /* EXB_C_* */
    int32_t arg;  /* EXB_C_SEQ_MAP: EXB_MAP_* */
} exb_computed;
#define EXB_MAX_COMPUTED 8

#define EXB_RD_STRING_T 1 /* VARCHAR columns also come back as arrays of 16-byte DuckDB string_t (string_type.hpp:19-60):
                             {u32 length, 12 inlined bytes} or {u32 length, 4-byte prefix, pointer into the batch's
                             host buffer}, built on the device -- the host's SetVectorString loop
                             (arrow_conversion.cpp:252-266) becomes one pointer assignment per column           */
#define EXB_RD_NO_OFFSETS 2 /* do not copy the Arrow-style int64 offsets back (hosts that only read string_t).  With both
                               flags a column whose rows are byte ranges of the input (every FASTQ column, FASTA id /
                               description) may come back with offsets == data == NULL: its string_t pointers then
                               lead into the pinned input block the reader staged the file in -- nothing was gathered
                               or copied back -- and stay valid until exb_batch_release like every other pointer     */

#define EXB_RD_COPY_IO 4    /* always stage the file through pinned blocks (page cache -> pinned memory -> device) and do not
                               have its page cache registered: the path a FIRST scan of a file takes anyway.  Without the
                               flag, a plain file that was scanned once is registered with CUDA as soon as the process's readers pause (its
                               page-cache pages become pinned memory; memory-backed file systems such as tmpfs, the file
                               must be writable by the process although nothing is written) and every later scan copies
                               it to the device by DMA straight from there.  exb_file_cache_state tells which applies.  */

typedef struct exb_reader_options {
    uint32_t size;       /* sizeof(exb_reader_options) of the caller (versioning)                                  */
    int32_t device;      /* CUDA device ordinal that runs this reader's pipeline; -1 = the caller's current device */
    /* Byte-range shard of ONE uncompressed file: this reader owns the records whose first byte lies in
     * [range_lo, range_hi).  Both ends are moved forward to the next record start by the same host-side rule
     * (FASTA: next line that starts with '>'; FASTQ: next line start where '@' / '+' alternate two lines apart over a
     * window of records), so adjacent shards agree on the cut by construction; the shard is then parsed as a complete
     * input, and a FASTQ cut that is not a true record boundary fails the scan's own line-count / '@' / '+' checks --
     * an error, never a silently different result.  range_hi <= 0: no upper limit.                               */
    int64_t range_lo, range_hi;
    /* Directory inputs: of the sorted file list, this reader takes the files [file_lo, file_hi); file_hi <= 0: all. */
    int32_t file_lo, file_hi;
    uint32_t column_mask; /* bit c: file column c is wanted                                                        */
    uint32_t flags;       /* EXB_RD_*                                                                              */
    int32_t n_computed;
    exb_computed computed[EXB_MAX_COMPUTED];
} exb_reader_options;

#define EXB_T_VARCHAR 0
#define EXB_T_FLOAT 1
#define EXB_T_DOUBLE 2
#define EXB_T_INT32_LIST 3
#define EXB_T_INT64 4
typedef struct exb_column_view {
    const int64_t *offsets; /* VARCHAR: n_rows + 1 entries (NOT rebased to 0), NULL if projected out / not requested */
    const uint8_t *data;    /* VARCHAR: the bytes `offsets` index                                                     */
    const uint8_t *valid;   /* one byte per row (0 = NULL), or NULL = all rows valid                                  */
    int32_t type;           /* EXB_T_*                                                                                */
    int32_t kind;           /* 0 = column of the file, else EXB_C_*                                                   */
    const void *strings;    /* VARCHAR with EXB_RD_STRING_T: n_rows 16-byte string_t entries                          */
    int64_t chunk_nulls;    /* nullable columns: NULL rows in the chunk this batch is a slice of (0: skip `valid`)        */
    const uint64_t *valid_bits; /* the same validity as a bitmap (bit i%64 of word i/64) when chunk_nulls > 0 and the
                                   batch starts at a multiple of 64 rows of its chunk (batch_rows % 64 == 0), else NULL */
    const void *values;     /* FLOAT / DOUBLE / INT64: n_rows values;  INT32_LIST: the n_values child values          */
    const uint64_t *list_entries; /* INT32_LIST: n_rows x {offset, length} (DuckDB list_entry_t), offsets into values */
    int64_t n_values;
} exb_column_view;
typedef struct exb_batch {
    int64_t n_rows;         /* 0 = end of stream */
    int32_t n_cols;         /* columns of the file; the computed columns follow them in cols[] */
    int32_t n_computed;
    int64_t batch_index;    /* position of this batch in the reader's output order (0, 1, 2, ...) */
    exb_column_view cols[4 + EXB_MAX_COMPUTED];
    void *owner;
} exb_batch;
EXB_API int exb_reader_open(const char *uri, const char *file_format, const char *compression, int64_t batch_rows,
                            const char *filters, uint32_t column_mask, exb_reader **out);
/* The same with a device, a shard and computed columns (options->column_mask replaces the argument). */
EXB_API int exb_reader_open2(const char *uri, const char *file_format, const char *compression, int64_t batch_rows,
                             const char *filters, const exb_reader_options *options, exb_reader **out);
/* column names in schema order; returns the number of columns */
EXB_API int exb_reader_columns(const exb_reader *reader, const char **names, int cap);
/* Thread-safe: several host threads may pull batches from one reader (each batch goes to exactly one of them). */
EXB_API int exb_reader_next(exb_reader *reader, exb_batch *out);
EXB_API void exb_batch_release(exb_batch *batch);
/* COUNT(*): consumes the rest of the stream (rows of a batch that was already handed out are not counted again);
 * rows that pass the filters are counted on the device and nothing is gathered or copied back
 * (arrow_conversion.cpp:813-816 is the reference's row-id-only case). */
EXB_API int exb_reader_count(exb_reader *reader, int64_t *n_rows);
/* Input bytes handed to the device so far / input bytes in total (table_scan_progress). */
EXB_API int exb_reader_progress(const exb_reader *reader, int64_t *bytes_done, int64_t *bytes_total);
/* Size in bytes of the (uncompressed, single-file) input a uri names, the number of files of a directory input, and
 * whether byte-range shards apply (1) -- what a host needs to plan one reader per GPU. */
EXB_API int exb_reader_plan(const char *uri, const char *file_format, const char *compression, int64_t *total_bytes,
                            int32_t *n_files, int32_t *range_shardable);
EXB_API void exb_reader_close(exb_reader *reader);
/* Registered page cache of a plain input file: 0 = not registered (never scanned, not possible, or switched off),
 * 1 = registration running in the background, 2 = registered: the next scan reads it by DMA from the page cache,
 * negative = the file cannot be inspected. */
EXB_API int exb_file_cache_state(const char *path);
/* 1 if the file this reader is reading now comes by DMA from its registered page cache, 0 if it is staged through
 * pinned blocks. */
EXB_API int exb_reader_io_path(const exb_reader *reader);
/* Per-scan counters (SURVEY 5 "exon_gpu_stats()"): every reader that ran leaves one record when it is closed; the most
 * recent 64 are kept, newest first.  The extension serves them as the table function exon_gpu_stats(). */
typedef struct exb_scan_stats {
    char path[256];        /* first input file                                                            */
    int32_t n_files;
    int32_t format;        /* 1 = FASTA, 2 = FASTQ                                                        */
    int32_t device;
    int32_t compression;   /* 0 none, 1 gzip (zlib stream), 2 zstd, 3 bzip2, 4 xz, 5 BGZF inflated on the device */
    int32_t io_path;       /* 1 = DMA from the registered page cache, 0 = copied through pinned blocks    */
    int32_t failed;
    int64_t file_bytes, bytes_done, rows, blocks;
    double seconds_total, seconds_io, seconds_device, seconds_scan, seconds_select, seconds_materialise,
        seconds_first_block;
} exb_scan_stats;
EXB_API int exb_stats_snapshot(exb_scan_stats *out, int max, int *n);
EXB_API void exb_stats_reset(void);
/* CUDA devices visible to the process (0 when there is none). */
EXB_API int exb_device_count(void);

/* ---------------------------------------------------------------------------
 * (2) Device layer
 * ------------------------------------------------------------------------- */

/* what a FASTQ scan writes */
#define EXB_F_LINES 1 /* line_end[g]: offset of the '\n' that ends line g       */
#define EXB_F_SEQ 2   /* seq_len[r], gc[r]                                       */
#define EXB_F_QUAL 4  /* qual_len[r], qsum[r] = sum((signed char)c - 33)         */
#define EXB_F_LOCAL_RECORDS 16 /* a chained / sharded range indexes its per-record outputs from the FIRST record it touches
                                  (record of global line prev.total_lines) instead of from record 0 of the file: what a
                                  byte-range shard needs, whose arrays only cover its own records.  A record that
                                  straddles the range's first byte gets index 0 and only the fields whose line ENDS in
                                  the range; zero the arrays first if partial records matter                          */

typedef struct exb_scan_result {
    uint64_t total_lines;     /* FASTQ: lines seen, the unterminated last line included  */
    int64_t open_line_start;  /* offset of the first byte after the last newline         */
    uint64_t err_pos;         /* smallest offset of a malformed line start; ~0 = none    */
    uint32_t overflow;        /* 1 = an output capacity (or the workspace) was too small */
    uint32_t pad;             /* FASTQ chaining: first byte of the open line is '@' (2) / '+' (1) */
    uint64_t n_records;       /* FASTA: number of '>' header lines                       */
    uint64_t seq_bytes;       /* FASTA: sequence bytes kept (line terminators stripped)  */
    uint64_t gc_total;        /* FASTA: G/C among them                                   */
    int64_t tail_s, tail_g;   /* FASTQ: state of the open line (used when chaining chunks) */
    uint64_t tail_hdr;        /* FASTA: the open line is a header line (chaining)          */
    uint64_t crlf_lines;      /* FASTQ general scan: lines of the range that ended in CR LF (0 lets
                                 exb_fastq_fields skip its three CR probes per record)       */
} exb_scan_result;

/* Bytes of scratch a scan over n input bytes needs (the calls clear what they need cleared).
 * For a FASTQ scan the workspace also holds one 8-byte record per line between the
 * two kernels of the scan; this size assumes lines of >= 24 bytes on average.  A
 * scan that meets more lines reports overflow = 1: retry with a workspace of
 * exb_fastq_workspace_bytes(n, max_lines) (max_lines = n + 1 always suffices). */
EXB_API int64_t exb_scan_workspace_bytes(int64_t n);
EXB_API int64_t exb_fastq_workspace_bytes(int64_t n, int64_t max_lines);

/*
 * FASTQ scan of d_buf[begin, n) (noodles-fastq read_record semantics):
 * strict 4-line records, LF or CRLF, last line may be unterminated.
 * Chunked input: scan consecutive ranges of ONE device buffer in stream order,
 * passing the previous call's workspace as d_prev_workspace (NULL for the first
 * range, whose `begin` must be a record start) and is_final = 1 only for the
 * range that ends the input; line / record indices and the outputs they address
 * continue across the calls.  The previous workspace must stay untouched until
 * this call has completed on `stream`.  Lines with index >= max_lines are ignored
 * (UINT64_MAX = all).  d_buf must be 16-byte aligned and `begin` of a chained
 * range a multiple of 16.  d_workspace: exb_scan_workspace_bytes(n)
 * bytes, 16-byte aligned.  The scalar results land in
 * the first sizeof(exb_scan_result) bytes of the workspace; fetch them with
 * exb_scan_result_fetch.
 *   d_line_end : uint32_t[line_cap] (wide_offsets = 0, n < 4 GiB) or uint64_t[line_cap]
 *   d_seq_len, d_gc, d_qual_len : uint32_t[rec_cap];  d_qsum : int32_t[rec_cap]
 */
EXB_API int exb_fastq_scan(const void *d_buf, int64_t begin, int64_t n, int is_final,
                           const void *d_prev_workspace, uint64_t max_lines, int flags, void *d_line_end, int64_t line_cap, int wide_offsets, uint32_t *d_seq_len,
                           uint32_t *d_gc, uint32_t *d_qual_len, int32_t *d_qsum, int64_t rec_cap,
                           void *d_workspace, int64_t workspace_bytes, void *stream);

/* Synchronises `stream`, copies the result block out of the workspace. */
EXB_API int exb_scan_result_fetch(const void *d_workspace, exb_scan_result *out, void *stream);

/* Predicates evaluated per record on the scan's outputs, ANDed together. */
#define EXB_P_MEAN_QUALITY 0 /* list_avg(quality_score_string_to_list(q)) : DOUBLE, NULL when empty */
#define EXB_P_GC_CONTENT 1   /* gc_content(sequence) : FLOAT widened to DOUBLE                      */
#define EXB_P_SEQ_LEN 2      /* length(sequence)                                                     */
#define EXB_P_QUAL_LEN 3     /* length(quality_scores)                                               */
#define EXB_OP_GT 0
#define EXB_OP_GE 1
#define EXB_OP_LT 2
#define EXB_OP_LE 3
#define EXB_OP_EQ 4
#define EXB_OP_NE 5
typedef struct exb_predicate {
    int32_t field;
    int32_t op;
    double value;
} exb_predicate;
#define EXB_MAX_PREDICATES 8

/* d_agg (int64_t[8], zeroed by the call): over the PASSING records
 *   [0] count  [1] sum seq_len  [2] sum gc  [3] sum qsum  [4] sum qual_len
 * d_pass (optional): uint8_t[n_records] 1 = passes.  Arrays a predicate or an
 * aggregate does not need may be NULL.  d_scan_workspace (optional): the
 * workspace of the exb_fastq_scan that produced the arrays; the kernel then
 * takes min(n_records, total_lines / 4) from it, so scan + filter can be
 * enqueued back to back without a host round trip. */
EXB_API int exb_fastq_filter(const uint32_t *d_seq_len, const uint32_t *d_gc, const uint32_t *d_qual_len,
                             const int32_t *d_qsum, int64_t n_records, const exb_predicate *preds, int n_preds,
                             uint8_t *d_pass, int64_t *d_agg, const void *d_scan_workspace, void *stream);

/*
 * Scan and filter in one kernel, for COUNT-style queries whose predicates only
 * read the quality line (EXB_P_MEAN_QUALITY, EXB_P_QUAL_LEN) -- the C2 query
 *   SELECT COUNT(*) FROM read_fastq(f) WHERE list_avg(quality_score_string_to_list(quality_scores)) > c
 * Same scan as exb_fastq_scan (same chaining rules, same result block: check
 * err_pos and total_lines % 4 after exb_scan_result_fetch), but each quality
 * line is tested as it is emitted and nothing per record is written:
 *   d_agg[0] = passing records, d_agg[3] = sum qsum, d_agg[4] = sum qual_len over
 *   them (the other entries stay 0).  accumulate = 0 clears d_agg (int64_t[8])
 *   first; 1 adds to it (later ranges of a chained scan).
 */
EXB_API int exb_fastq_scan_filter(const void *d_buf, int64_t begin, int64_t n, int is_final, const void *d_prev_workspace,
                                  const exb_predicate *preds, int n_preds, int64_t *d_agg, int accumulate,
                                  void *d_workspace, int64_t workspace_bytes, void *stream);

/*
 * Byte-range shards and any other case where the state BEFORE a range is only known after the range
 * has been read: K1 of the scan (the byte pass) never depends on that state, so
 *   1. run exb_fastq_scan / exb_fastq_scan_filter on the range with a PROVISIONAL predecessor (a 128-byte device
 *      block written by exb_scan_result_store from a zeroed exb_scan_result whose open_line_start = begin), fetch the
 *      result: total_lines = newlines in the range, open_line_start / tail_s / tail_g / pad = its open last line;
 *   2. exchange those blocks (one all-gather of 88 bytes per shard), compose the true predecessor state;
 *   3. call the matching *_resolve function with the same begin / n / outputs / workspace and the true predecessor:
 *      only the light second kernel runs again (8 bytes per line, or 48 bytes per tile for the fused flavour); the
 *      input bytes are not read a second time.  d_agg / per-record outputs of step 1 are overwritten.
 * exon_duckdb_b200/dist.py is the host side of this protocol.
 */
/* Fused TOTALS flavour (BASELINE config C5: SELECT COUNT(*), SUM(length(sequence)), SUM(#GC), AVG(gc_content(sequence))):
 * the byte pass adds up, per 4 KiB tile and per phase hypothesis, the aggregates of every line that would be a
 * SEQUENCE line; K2 picks the bucket of the tile's true phase.  Nothing per record or per line is written.
 *   d_agg int64[8]: [1] sum of sequence lengths  [2] sum of G/C  [3] sequence lines  [5] sum of round(gc_content * 2^32)
 *   with gc_content = (float)#GC / (float)len per record as the scalar function defines it (0 for an empty sequence);
 *   the record count is total_lines / 4 of the result block.  Same layout as exb_fastq_seq_totals over the per-record
 *   arrays of the general scan, and bit-identical to it.  accumulate != 0 adds to d_agg instead of zeroing it first.
 * _begin / _resolve split the call for byte-range shards exactly like exb_fastq_scan_begin / _resolve: K1 + line
 * offsets + result block first, K2 under the true predecessor (exb_fastq_compose_prev) after the block exchange. */
EXB_API int exb_fastq_scan_totals(const void *d_buf, int64_t begin, int64_t n, int is_final, const void *d_prev_workspace,
                                  int64_t *d_agg, int accumulate, void *d_workspace, int64_t workspace_bytes, void *stream);
EXB_API int exb_fastq_scan_totals_begin(const void *d_buf, int64_t begin, int64_t n, int is_final,
                                        const void *d_prev_workspace, void *d_workspace, int64_t workspace_bytes,
                                        void *stream);
EXB_API int exb_fastq_scan_totals_resolve(int64_t begin, int64_t n, int is_final, const void *d_prev_workspace,
                                          int64_t *d_agg, int accumulate, void *d_workspace, int64_t workspace_bytes,
                                          void *stream);

/* Step 1 of the fused COUNT flavour without its K2: the byte pass, the line-offset scan and the result block only
 * (a shard does not know its phase yet, so a provisional K2 would be thrown away).  Follow with the exchange and
 * exb_fastq_scan_filter_resolve on EVERY shard (the first one with d_prev_workspace = NULL). */
EXB_API int exb_fastq_scan_filter_begin(const void *d_buf, int64_t begin, int64_t n, int is_final,
                                        const void *d_prev_workspace, const exb_predicate *preds, int n_preds,
                                        void *d_workspace, int64_t workspace_bytes, void *stream);
/* Step 1 of the general flavour without its K2 (see exb_fastq_scan_filter_begin): byte pass, line-offset scan and
 * the result block.  flags = the EXB_F_* the later exb_fastq_scan_resolve will ask for. */
EXB_API int exb_fastq_scan_begin(const void *d_buf, int64_t begin, int64_t n, int is_final, const void *d_prev_workspace,
                                 int flags, void *d_workspace, int64_t workspace_bytes, void *stream);
EXB_API int exb_fastq_scan_resolve(int64_t begin, int64_t n, int is_final, const void *d_prev_workspace, uint64_t max_lines,
                                   int flags, void *d_line_end, int64_t line_cap, int wide_offsets, uint32_t *d_seq_len,
                                   uint32_t *d_gc, uint32_t *d_qual_len, int32_t *d_qsum, int64_t rec_cap,
                                   void *d_workspace, int64_t workspace_bytes, void *stream);
EXB_API int exb_fastq_scan_filter_resolve(int64_t begin, int64_t n, int is_final, const void *d_prev_workspace,
                                          const exb_predicate *preds, int n_preds, int64_t *d_agg, int accumulate,
                                          void *d_workspace, int64_t workspace_bytes, void *stream);
/* Byte-range sharded COUNT with ONE exchange.  exb_fastq_scan_filter_candidates = byte pass + line offsets + a K2 that
 * runs for all four phases at once: d_record (256 bytes, int64[32]) receives the range's line count, its open last line,
 * four candidate aggregate sets (one per possible phase of the range's first line, that line excluded) and the description
 * of the first line itself -- the only line whose length / Phred sum / first byte depend on the predecessors.  The records
 * of all shards (in shard order, 256 bytes apart) + d_ranges (as for exb_fastq_compose_prev) give the aggregates of the
 * whole file on EVERY rank without a second collective: exb_fastq_combine_records locally (after any all-gather), or
 * exb_peer_count_fused = all-gather over NVLink peer memory + combine in one launch.
 *   d_total (int64[8]): [0] passing records [3] their Phred sums [4] their lengths [6] lines of the file mod 4
 *   [7] 1 = a line start contradicts its phase (malformed), -1 = a peer did not answer.
 * d_prev_workspace of the candidates call = the PROVISIONAL predecessor (open_line_start = begin), as for
 * exb_fastq_scan_filter_begin.  Replaces begin -> allgather -> compose -> resolve -> reduce (5 launches, 2 flag waits). */
EXB_API int exb_fastq_scan_filter_candidates(const void *d_buf, int64_t begin, int64_t n, int is_final,
                                             const void *d_prev_workspace, const exb_predicate *preds, int n_preds,
                                             void *d_record, void *d_workspace, int64_t workspace_bytes, void *stream);
EXB_API int exb_fastq_combine_records(const void *d_records, const int64_t *d_ranges, int world, const exb_predicate *preds,
                                      int n_preds, int64_t *d_total, void *stream);
/* Step 2 on the device (no host round trip): d_blocks = the result blocks of all `world` shards in shard order
 * (128 bytes each, e.g. the output of an NCCL all-gather of the first 128 bytes of each shard's workspace), d_ranges =
 * int64[world][3] {lo, hi, begin} of every shard (file offsets of its range; local offset of byte lo).  Writes the true
 * predecessor of shard `rank` (>= 1) to d_prev_out (128 bytes), ready to be passed as d_prev_workspace. */
EXB_API int exb_fastq_compose_prev(const void *d_blocks, const int64_t *d_ranges, int world, int rank, void *d_prev_out,
                                   void *stream);
/* ---- shard exchange over NVLink peer memory (exon_duckdb_b200/csrc/peer_exchange.cu) ----------------
 * What byte-range shards must agree on is one 128-byte scan result block per shard and 8 int64 aggregates.  Every
 * rank owns a symmetric buffer of exb_peer_bytes() bytes (zeroed, then a barrier, before step 1) that every peer
 * has mapped; d_peers = device array of `world` pointers to those buffers in rank order (torch symmetric memory:
 * handle.buffer_ptrs_dev).  `seq` = step number, 1, 2, ... identical on every rank.  world <= 16.
 *   exb_peer_allgather_block  stores d_block (128 B) into every peer's buffer and waits for every peer's block:
 *                             afterwards the `world` blocks sit, in rank order and 128 B apart, at
 *                             own_buffer + exb_peer_blocks_offset(seq) (the input of exb_fastq_compose_prev).
 *   exb_peer_count_reduce     d_total[0..8) = sum over ranks of {d_agg[0..6), lines of the file mod 4 (last shard),
 *                             1 if the shard's result block (first 128 B of d_workspace) reports a malformed record};
 *                             d_total[7] = -1 if a peer did not answer within 10 s.
 * Replaces ncclAllGather + ncclAllReduce of the same bytes (2 x ~150 us at 8 ranks) with two ~10 us launches. */
EXB_API int64_t exb_peer_bytes(void);
EXB_API int64_t exb_peer_blocks_offset(uint64_t seq);
EXB_API int exb_peer_allgather_block(void *const *d_peers, int rank, int world, const void *d_block, uint64_t seq,
                                     void *stream);
EXB_API int exb_peer_count_fused(void *const *d_peers, int rank, int world, const void *d_record, const int64_t *d_ranges,
                                 const exb_predicate *preds, int n_preds, uint64_t seq, int64_t *d_total, void *stream);
EXB_API int exb_peer_count_reduce(void *const *d_peers, int rank, int world, const void *d_workspace,
                                  const int64_t *d_agg, int is_last, uint64_t seq, int64_t *d_total, void *stream);

/* Writes *src where a scan expects a predecessor's result block (d_dst: >= 128 bytes of device memory). */
EXB_API int exb_scan_result_store(void *d_dst, const exb_scan_result *src, void *stream);

/* Field extents of FASTQ records from the line index: d_lens is uint32_t[4][n_records]
 * (name, description, sequence, quality_scores); d_desc_valid uint8_t[n_records]
 * (0 = NULL description).  d_sel (optional) lists the records to take; d_starts (optional)
 * int64_t[4][n_rows] receives the field offsets in d_buf (input of exb_gather_ranges).
 * d_scan_workspace (optional): the workspace of the exb_fastq_scan that produced d_line_end over the same range; when
 * its result block says no line ended in CR LF, the kernel skips the byte it would read before every line end. */
EXB_API int exb_fastq_fields(const void *d_buf, int64_t begin, int64_t n, const void *d_line_end, int wide_offsets,
                             const int64_t *d_sel, int64_t n_rows, uint32_t *d_lens, uint8_t *d_desc_valid,
                             int64_t *d_starts, const void *d_scan_workspace, void *stream);

/* Exclusive prefix sum uint32 -> int64 with the total in d_out[n] (n + 1 outputs).
 * d_workspace: exb_scan_workspace_bytes(4 * n) zero-initialised by the call. */
EXB_API int exb_exclusive_scan_u32(const uint32_t *d_in, int64_t n, int64_t *d_out, void *d_workspace,
                                   int64_t workspace_bytes, void *stream);
/* The same for `cols` (<= 8) independent columns in ONE launch: column c reads d_in + c * in_stride and writes
 * d_out + c * out_stride (n + 1 entries each).  The four field-length columns of exb_fastq_fields become the
 * four Arrow offset arrays of a read_fastq batch this way.  d_workspace: exb_scan_workspace_bytes(4 * n * cols). */
EXB_API int exb_exclusive_scan_u32_multi(const uint32_t *d_in, int64_t n, int cols, int64_t in_stride, int64_t *d_out,
                                         int64_t out_stride, void *d_workspace, int64_t workspace_bytes, void *stream);

/* Row ids of the records whose d_pass byte is set: d_sel[0..count), count -> d_offsets[n]
 * where d_offsets (int64[n+1]) is scratch. */
EXB_API int exb_select_rows(const uint8_t *d_pass, int64_t n, int64_t *d_offsets, int64_t *d_sel, void *d_workspace,
                            int64_t workspace_bytes, void *stream);

/* Copy column `col` (0..3) of the selected FASTQ records into d_out at d_off[row]. */
EXB_API int exb_fastq_gather(const void *d_buf, int64_t begin, int64_t n, const void *d_line_end, int wide_offsets,
                             const int64_t *d_sel, int64_t n_rows, int col, const uint32_t *d_lens,
                             const int64_t *d_off, uint8_t *d_out, void *stream);
/* The field split, the offset scan and the four column gathers of ALL records of a scanned range as ONE launch
 * (record_ops.cu fastq_split_kernel): a block of 256 consecutive records computes its field extents, obtains its output
 * offsets with a decoupled look-back over the earlier blocks and copies its rows of every wanted column while their
 * input window is hot in L2, so each input DRAM atom is fetched once.  Replaces exb_fastq_fields +
 * exb_exclusive_scan_u32_multi + 4 x exb_fastq_gather for unfiltered scans.
 *   column_mask : bit c = copy column c (name, description, sequence, quality_scores); offsets are produced for all four
 *   d_off       : int64[4][n_rows + 1] Arrow offsets (exclusive prefix of the field lengths), column-major
 *   d_out[c]    : 16-byte aligned output of column c, cap[c] bytes (NULL / 0 for columns not in the mask)
 *   d_scratch   : exb_fastq_split_scratch_bytes(n_rows) bytes (zeroed by the call); after the call its word [2] (uint32)
 *                 is 1 if a column did not fit its capacity
 *   map_mode    : < 0 = plain copy; else EXB_MAP_* applied to the sequence column on its way out, *d_bad as in
 *                 exb_fastq_gather_map */
EXB_API int64_t exb_fastq_split_scratch_bytes(int64_t n_rows);
EXB_API int exb_fastq_split(const void *d_buf, int64_t begin, int64_t n, const void *d_line_end, int wide_offsets,
                            int64_t n_rows, uint32_t column_mask, int64_t *d_off, uint8_t *d_desc_valid,
                            uint8_t *const *d_out, const int64_t *cap, void *d_scratch, const void *d_scan_workspace,
                            int map_mode, uint64_t *d_bad, void *stream);
/* The same gather with reverse_complement / complement (sequence_functions/module.cpp:30-121, the reference's LUTs)
 * applied to the bytes on their way out: `SELECT reverse_complement(sequence) FROM read_fastq(...)` on a
 * device-resident file is one pass over the sequence bytes instead of gather + exb_seq_map.  mode = EXB_MAP_*.
 * *d_bad = ~0 if every byte was one of ACGT, else (position in d_out << 8) | offending byte, smallest position. */
EXB_API int exb_fastq_gather_map(const void *d_buf, int64_t begin, int64_t n, const void *d_line_end, int wide_offsets,
                                 const int64_t *d_sel, int64_t n_rows, int col, const uint32_t *d_lens,
                                 const int64_t *d_off, int mode, uint8_t *d_out, uint64_t *d_bad, void *stream);


/*
 * FASTA scan of d_buf[begin, n) (noodles-fasta read_definition / read_sequence
 * semantics): records start at a line whose first byte is '>'; the sequence is
 * the concatenation of the following lines with LF / CRLF removed.  Per record
 * r: d_hdr_start[r] = offset of '>', d_hdr_end[r] = offset of the header
 * line's terminator; d_seq_off / d_gc_prefix (rec_cap + 1 entries) = sequence
 * bytes / G,C bytes before record r, closed with the totals.  d_seq_out
 * (optional, seq_cap bytes) receives the compacted sequence column.
 * Chunk chaining as for exb_fastq_scan; halo_n (>= n) = how many bytes of d_buf
 * are already valid: a CR in the last byte of a non-final range needs one byte
 * of look-ahead to know whether an LF follows.
 */
EXB_API int exb_fasta_scan(const void *d_buf, int64_t begin, int64_t n, int is_final, int64_t halo_n,
                           const void *d_prev_workspace, int64_t *d_hdr_start, int64_t *d_hdr_end,
                           int64_t *d_seq_off, int64_t *d_gc_prefix, int64_t rec_cap, uint8_t *d_seq_out,
                           int64_t seq_cap, void *d_workspace, int64_t workspace_bytes, void *stream);

/* id / description extents of FASTA header lines (Definition::from_str):
 * d_lens uint32_t[2][n_rows], d_starts int64_t[2][n_rows] (offsets of id / description in
 * d_buf, ready for exb_gather_ranges), d_desc_valid uint8_t[n_rows];
 * *d_err_pos (set to ~0 first) = smallest offset of a header with an empty name. */
EXB_API int exb_fasta_headers(const void *d_buf, int64_t n, const int64_t *d_hdr_start, const int64_t *d_hdr_end,
                              int64_t n_rows, uint32_t *d_lens, int64_t *d_starts, uint8_t *d_desc_valid,
                              uint64_t *d_err_pos, void *stream);

/* Generic gather of byte ranges: d_out[d_off[i] .. d_off[i+1]) = d_buf[d_start[i] ..).  d_off = the exclusive prefix sum
 * of the row lengths, n_rows + 1 entries (exb_exclusive_scan_u32); out_bytes = size of d_out (>= d_off[n_rows]).
 * d_len is unused (kept for symmetry with the other calls; may be NULL). */
EXB_API int exb_gather_ranges(const void *d_buf, const int64_t *d_start, const uint32_t *d_len,
                              const int64_t *d_off, int64_t n_rows, uint8_t *d_out, int64_t out_bytes, void *stream);

/* gc_content per record from FASTA scan prefixes: out[r] = (float)gc / (float)len, '' -> 0. */
EXB_API int exb_gc_from_prefix(const int64_t *d_seq_off, const int64_t *d_gc_prefix, int64_t n_rows, float *d_out,
                               void *stream);
/* same from FASTQ per-record counts */
EXB_API int exb_gc_from_counts(const uint32_t *d_seq_len, const uint32_t *d_gc, int64_t n_rows, float *d_out,
                               void *stream);

/* Totals over the per-record arrays of a scan, for COUNT / SUM / AVG queries (BASELINE config C5:
 * SELECT COUNT(*), SUM(#GC), SUM(length(sequence)), AVG(gc_content(sequence))):  d_totals (int64[8], ADDED to)
 *   [1] += sum seq_len   [2] += sum gc   [5] += sum round(gc_content(r) * 2^32)  (fixed point: the float average is
 *   then independent of the summation order and of the number of shards; |error| <= 2^-33 per record)
 * Entries are additive per sequence LINE, so byte-range shards may simply add theirs (EXB_F_LOCAL_RECORDS arrays). */
EXB_API int exb_fastq_seq_totals(const uint32_t *d_seq_len, const uint32_t *d_gc, int64_t n_records, int64_t *d_totals,
                                 void *stream);

/* ---- scalar functions over an Arrow-style string column (int64 offsets + bytes) ----
 * sequence_functions/module.cpp:131-158 (per-row formula; d_valid NULL = all valid;
 * invalid rows produce 0 and stay invalid in the caller's mask). */
EXB_API int exb_gc_content(const int64_t *d_off, const uint8_t *d_data, int64_t n_rows, float *d_out, void *stream);

#define EXB_MAP_REVERSE_COMPLEMENT 0 /* module.cpp:30-69  A->C T->G C->A G->T (reference semantics) */
#define EXB_MAP_COMPLEMENT 1         /* module.cpp:81-121 A<->T C<->G                                 */
#define EXB_MAP_TRANSCRIBE 2         /* module.cpp:212-249 T->U, A C G unchanged                      */
#define EXB_MAP_REVERSE_TRANSCRIBE 3 /* module.cpp:168-203 U->T, A C G unchanged                      */
/* Maps n_bytes bytes through the table; *d_bad_pos (uint64, set to ~0 first)
 * receives the smallest index of a byte outside ACGT (the reference throws
 * InvalidInputException for it). */
EXB_API int exb_seq_map(const uint8_t *d_in, int64_t n_bytes, int mode, uint8_t *d_out, uint64_t *d_bad_pos,
                        void *stream);

/* translate_dna_to_aa (module.cpp:260-360) over a string column: standard codon table over upper-case ACGT.
 * d_off = int64[n_rows + 1] offsets into d_data (offsets keep their base, as everywhere).  A row whose length is not a
 * multiple of 3 is the reference's "Invalid sequence length: <len>"; a codon outside the table "Invalid codon: <xyz>".
 * Rows are checked in order and a row's length before its codons, so d_status (int64[2], device) receives
 *   [0] the first row with a bad length, or -1;   [1] the data offset of the first bad codon BEFORE that row, or -1.
 * When both are -1, d_out holds (d_off[n_rows] - d_off[0]) / 3 bytes and row i's protein is
 * d_out[(d_off[i] - d_off[0]) / 3 .. (d_off[i+1] - d_off[0]) / 3). */
EXB_API int exb_translate(const int64_t *d_off, const uint8_t *d_data, int64_t n_rows, uint8_t *d_out,
                          int64_t *d_status, void *stream);

/* fastq_functions/module.cpp:32-50: out[i] = (signed char)in[i] - 33 (list child vector;
 * the list offsets are the string offsets). */
EXB_API int exb_quality_decode(const uint8_t *d_in, int64_t n_bytes, int32_t *d_out, void *stream);

/* ---- the same scalar functions over HOST columns: what a DuckDB scalar-function callback holds for one
 * DataChunk (<= 2048 rows).  H2D copy, kernel, D2H copy on a stream owned by the calling thread; synchronous.
 * Callers: exon_duckdb_b200/duckdb_ext (gc_content, reverse_complement, complement, quality_score_string_to_list). */
EXB_API int exb_gc_content_host(const int64_t *offsets, const uint8_t *data, int64_t n_rows, float *out);
/* *bad_pos = -1, or the smallest index of a byte outside ACGT (out is then unspecified) */
EXB_API int exb_seq_map_host(const uint8_t *data, int64_t n_bytes, int mode, uint8_t *out, int64_t *bad_pos);
/* offsets: n_rows + 1 entries starting at 0 into `data`; out: offsets[n_rows] / 3 bytes; status[2] as exb_translate
 * (status[1] = offset into `data`). */
EXB_API int exb_translate_host(const int64_t *offsets, const uint8_t *data, int64_t n_rows, uint8_t *out,
                               int64_t *status);
EXB_API int exb_quality_decode_host(const uint8_t *data, int64_t n_bytes, int32_t *out);

/* ---- writers: the inverse of the scan --------------------------------------------------------
 * Replace FastqFunctions::GetFastqCopyFunction (exon/include/exon/fastq_functions/module.hpp:30) and its FASTA twin:
 * `COPY ... TO 'file' (FORMAT 'fastq' | 'fasta' [, COMPRESSION 'gzip' | 'zstd'] [, FORCE true])`
 * (test/sql/exondb-release-with-deb-info/test_fastq_copy.test, test_fasta_copy.test; both commented out in the
 * reference because it removed the writers).  Record layout (noodles-fastq / noodles-fasta writers):
 *   FASTQ  '@' name [' ' description] '\n' sequence '\n' '+' '\n' quality_scores '\n'
 *   FASTA  '>' id [' ' description] '\n' sequence in lines of `line_width` bases (noodles default: 80), each + '\n'
 * A NULL or empty description writes no space.
 *
 * Device layer: Arrow-style columns already in HBM -> file image in HBM.  d_off[c] has n_rows + 1 entries (need not
 * start at 0) and indexes d_data[c] (4-byte aligned base); columns are name / id, description, sequence and -- FASTQ
 * only -- quality_scores; d_desc_valid (one byte per row) may be NULL.  d_row_off receives n_rows + 1 record offsets
 * of the image.  d_scratch: exb_format_scratch_bytes(n_rows).  Asynchronous on `stream`; exb_format_finish
 * synchronises, reports the image size and fails with EXB_ERR_CAPACITY if out_cap was too small (records that did
 * not fit were skipped, nothing was written out of bounds). */
typedef struct exb_format_cols {
    const int64_t *d_off[4];
    const uint8_t *d_data[4];
    const uint8_t *d_desc_valid;
} exb_format_cols;
EXB_API int64_t exb_format_scratch_bytes(int64_t n_rows);
EXB_API int exb_fastq_format(const exb_format_cols *cols, int64_t n_rows, int64_t *d_row_off, uint8_t *d_out,
                             int64_t out_cap, void *d_scratch, int64_t scratch_bytes, void *stream);
EXB_API int exb_fasta_format(const exb_format_cols *cols, int64_t n_rows, int line_width, int64_t *d_row_off,
                             uint8_t *d_out, int64_t out_cap, void *d_scratch, int64_t scratch_bytes, void *stream);
EXB_API int exb_format_finish(const int64_t *d_row_off, int64_t n_rows, const void *d_scratch, int64_t out_cap,
                              int64_t *out_bytes, void *stream);

/* Host layer: a file being written.  compression: NULL / "" / "auto" = by suffix (.gz, .zst), "gzip" | "gz",
 * "zstd" | "zst", "none"; force = 0 refuses to replace an existing regular file (the reference's FORCE option).
 * exb_writer_append takes HOST columns in the same offsets + bytes form (offsets need not start at 0), stages them in
 * pinned memory and formats on `device` whenever ~64 MiB are staged; rows keep their order.  exb_writer_close flushes,
 * closes the file and frees the writer (also after a failure). */
typedef struct exb_writer exb_writer;
EXB_API int exb_writer_open(const char *path, const char *file_format, const char *compression, int force, int device,
                            exb_writer **out);
EXB_API int exb_writer_set_line_width(exb_writer *writer, int line_width); /* FASTA; default 80 */
EXB_API int exb_writer_append(exb_writer *writer, int64_t n_rows, const int64_t *const *offsets,
                              const uint8_t *const *data, const uint8_t *desc_valid);
EXB_API int exb_writer_close(exb_writer *writer, int64_t *rows_written, int64_t *bytes_written);

/* ---- BGZF (blocked gzip) input inflated on the device: SURVEY 8(f) rank 1 -------------------------------------
 * Replaces, for bgzip'ed files, the streaming GzipDecoder the reference puts in front of its parser
 * (rust/src/arrow_reader.rs:60-91, datafusion FileCompressionType::GZIP).  A BGZF file is a chain of independent gzip
 * members of <= 64 KiB (SAM specification 4.1).  The host walks the member headers -- no decompression --, the
 * compressed bytes go to the device as they are, one warp inflates one member (DEFLATE, RFC 1951) and checks its
 * CRC-32.  The reader does this by itself for .gz inputs that are BGZF (exb_reader_open*); these are its pieces. */
typedef struct exb_bgzf_block {
    int64_t in_off;   /* offset of the member's DEFLATE payload, relative to the first byte handed to the device */
    int64_t out_off;  /* offset of its text in the output                                                        */
    uint32_t clen;    /* payload bytes                                                                           */
    uint32_t isize;   /* text bytes (gzip ISIZE)                                                                 */
    uint32_t crc32;   /* CRC-32 of the text (gzip trailer)                                                       */
    uint32_t reserved;
} exb_bgzf_block;
/* 1 if `bytes` (HOST memory, the first n bytes of a file) start with a BGZF member header. */
EXB_API int exb_bgzf_probe_host(const uint8_t *bytes, int64_t n);
/* HOST: list the members that start at compressed offset `pos` of the n-byte compressed image `bytes` until their text
 * would exceed max_out_bytes (at least one member is taken) or max_blocks are listed.  in_off is relative to `pos`.
 * next_pos = compressed offset of the first member not listed; out_bytes = text bytes of the listed ones. */
EXB_API int exb_bgzf_index_host(const uint8_t *bytes, int64_t n, int64_t pos, int64_t max_out_bytes,
                                exb_bgzf_block *blocks, int64_t max_blocks, int64_t *n_blocks, int64_t *next_pos,
                                int64_t *out_bytes);
/* HOST: the inverse, for output -- `text` as BGZF members of 65280 text bytes (like htslib's bgzip), deflated by zlib at
 * `level` (0-9, else 6) on `threads` host threads; eof_marker != 0 appends the 28-byte empty member that ends a BGZF
 * file.  out_cap >= exb_bgzf_compress_bound(n).  What the writers' gzip sink produces (COPY ... TO, COMPRESSION 'gzip'):
 * any gzip reader reads it, and this library's reader inflates it on the device. */
EXB_API int64_t exb_bgzf_compress_bound(int64_t n);
EXB_API int exb_bgzf_compress_host(const uint8_t *text, int64_t n, int level, int threads, int eof_marker, uint8_t *out,
                                   int64_t out_cap, int64_t *out_bytes);
EXB_API int64_t exb_bgzf_scratch_bytes(void);
/* DEVICE: inflate the listed members of d_in (16 readable bytes of slack behind the last payload) into d_out.
 * Asynchronous on `stream`.  check_crc != 0 also verifies every member's CRC-32. */
EXB_API int exb_bgzf_inflate(const uint8_t *d_in, const exb_bgzf_block *d_blocks, int64_t n_blocks, uint8_t *d_out,
                             void *d_scratch, int check_crc, void *stream);
/* Synchronises `stream`; EXB_ERR_FORMAT (+ the index of the first bad member) if a member was corrupt. */
EXB_API int exb_bgzf_finish(const void *d_scratch, int64_t *bad_block, void *stream);
/* The same test on a HOST copy of the 16 scratch bytes (callers that read them back with their other results). */
EXB_API int exb_bgzf_status(const unsigned int *scratch_host, int64_t *bad_block);

/* ---- host-buffer engine (end-to-end path): parse a FASTQ held in host memory ----
 * Streams `n` host bytes through pinned staging buffers with double-buffered
 * cudaMemcpyAsync, runs scan + filter per chunk, returns the aggregates
 * (same layout as exb_fastq_filter's d_agg, plus [5] = records seen).
 * Synchronous.  device = CUDA device ordinal. */
EXB_API int exb_fastq_count_host(const void *host_buf, int64_t n, const exb_predicate *preds, int n_preds,
                                 int64_t chunk_bytes, int device, int64_t *agg_out, exb_scan_result *res_out);

/* The same with the device buffers, streams and events kept across calls. */
typedef struct exb_engine exb_engine;
EXB_API int exb_engine_create(int device, int64_t chunk_bytes, exb_engine **out);
EXB_API void exb_engine_destroy(exb_engine *engine);
EXB_API int exb_engine_fastq_count(exb_engine *engine, const void *host_buf, int64_t n, const exb_predicate *preds,
                                   int n_preds, int64_t *agg_out, exb_scan_result *res_out);
/* Page-locked host memory (cudaHostAlloc): the source layout the engine copies from at full PCIe rate. */
EXB_API void *exb_host_alloc(int64_t bytes);
EXB_API void exb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* EXON_B200_H */
