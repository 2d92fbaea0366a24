// exon_b200_internal.h -- kernel argument blocks shared between the .cu files.
// Nothing in here crosses the C ABI (see include/exon_b200.h for that).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/exon_b200.h"

// internal scan flag: evaluate quality-line predicates at emission, only aggregates leave the kernel
#define EXB_F_FUSED 8

namespace exb {

struct TileSlot;

// Device-resident scalar results of one scan launch.  Layout is mirrored by
// exb_scan_result in include/exon_b200.h (copied back verbatim).
struct ScanResult {
    uint64_t total_lines;       // FASTQ: newlines seen (incl. the virtual one at EOF)
    int64_t open_line_start;    // first byte after the last newline
    unsigned long long err_pos; // ON THE DEVICE: ~(smallest offset of a malformed line start), 0 = none (atomicMax over a
                                // zeroed word needs no init kernel); exb_scan_result_fetch inverts it for the host
    uint32_t overflow;          // an output capacity was too small
    uint32_t pad;               // FASTQ: first byte of the open line is '@' (bit 1) / '+' (bit 0) (chunk chaining)
    uint64_t n_records;         // FASTA: header lines seen
    uint64_t seq_bytes;         // FASTA: sequence bytes kept (newlines / CR stripped)
    uint64_t gc_total;          // FASTA: G/C among them
    int64_t tail_s, tail_g;     // FASTQ: byte sum / G,C count of the open line (chunk chaining)
    uint64_t tail_hdr;          // FASTA: the open line is a header line (chunk chaining)
    unsigned long long crlf_lines;  // FASTQ general scan (K2): lines that ended in CR LF
};

// per-tile aggregates of the fused COUNT flavour: one bucket per phase hypothesis h = (global index of the
// tile's first newline-terminated line) & 3.  Lines 1.. of the tile only; line 0 is completed by K2.
struct alignas(16) FusedTile {
    uint32_t cq[4];   // lines that are quality lines under h and pass the predicates | sum of their lengths << 12
    int32_t qs[4];    // sum of their Phred sums
    int32_t ps0;      // record of the tile's first newline (what K2 needs to finish line 0)
    uint32_t y0;
    uint32_t bad4;    // bit h: a line start contradicts h ('@' / '+' expected)
    uint32_t pad;
};
static_assert(sizeof(FusedTile) == 48, "FusedTile layout");

struct FastqScanArgs {
    const uint8_t* buf;
    int64_t begin, n;           // parse bytes [begin, n)
    const ScanResult* prev;     // null: `begin` starts line 0; else continue from that scan's final state
    int is_final;               // 1: n is the end of the input (an unterminated last line gets a virtual '\n')
    uint64_t max_lines;         // lines with index >= max_lines are ignored
    int64_t n_tiles;            // 4 KiB warp tiles
    int64_t tma_rows;           // 128-byte rows the tensor map covers (set by the launcher)
    ScanResult* result;
    // K1 -> K2 (all inside the caller's workspace; nothing needs zeroing except *rec_bump)
    uint32_t* tile_cnt;         // [n_tiles] newlines per tile
    int64_t* tile_rec;          // [n_tiles] index of the tile's first record
    int64_t* tile_rec2;         // [n_tiles] index of the second run of records | records in the first run << 48
    uint64_t* tails;            // [n_tiles] tail words
    const int64_t* line_base;   // [n_tiles + 1] exclusive scan of tile_cnt (written between K1 and K2)
    uint2* records;             // [rec_space] 8 bytes per newline, bump-allocated per warp
    int64_t rec_space;
    unsigned long long* rec_bump;
    FusedTile* fused_tiles;     // [n_tiles] (EXB_F_FUSED instead of records)
    int n_fused;                // EXB_F_FUSED: predicates on the quality line (EXB_P_MEAN_QUALITY / EXB_P_QUAL_LEN)
    exb_predicate fused[EXB_MAX_PREDICATES];
    long long* fused_agg;       // int64[8] aggregates (same layout as exb_fastq_filter's d_agg)
    unsigned long long* fused_fix;  // [n_tiles][4] EXB_F_FUSED | EXB_F_SEQ: sum of round(gc_content * 2^32) per hypothesis bucket
    // outputs
    void* line_end;             // OffT[line_cap]
    int64_t line_cap;
    uint32_t *seq_len, *gc, *qual_len;
    int32_t* qsum;
    int64_t rec_cap;
    int local_records;          // EXB_F_LOCAL_RECORDS: record indices count from the first record the range touches
};

struct FusedPreds {  // predicates by value (kernel argument)
    int n;
    exb_predicate p[EXB_MAX_PREDICATES];
};
cudaError_t fastq_candidates_launch(const FastqScanArgs& a, long long* rec, cudaStream_t st);
cudaError_t fastq_combine_records_launch(const long long* recs, const int64_t* ranges, int world, const exb_predicate* preds, int n_preds,
                                         long long* total, cudaStream_t st);
cudaError_t fastq_tile_launch(const FastqScanArgs& a, int flags, cudaStream_t st);                     // K1
cudaError_t fastq_emit_launch(const FastqScanArgs& a, int flags, bool wide_offsets, cudaStream_t st);  // K2
int64_t fastq_scan_tiles(int64_t begin, int64_t n, int is_final);  // 4 KiB warp tiles
int64_t fastq_record_slack(int64_t n_tiles);                       // record slots the bump allocation may waste

struct FaTile;
struct FastaScanArgs {
    const uint8_t* buf;
    int64_t begin, n;
    const ScanResult* prev;
    int is_final;
    int64_t halo_n;             // bytes of buf that are readable (>= n): the CRLF test looks one byte ahead
    int64_t n_tiles;            // 4 KiB warp tiles
    int64_t tma_rows;           // 128-byte rows the tensor map covers (set by the launcher)
    void* payload;              // workspace behind the header: fasta_workspace_payload(n_tiles) bytes, carved by the launcher
    FaTile* tiles;              // [n_tiles] K1's summaries
    uint8_t* tile_state;        // [n_tiles] state at the tile's first byte | 4 = the tile has per-record outputs
    int64_t* tile_base3;        // [n_tiles][3] records / kept bytes / G,C before the tile
    ScanResult* result;
    // per record r (single writer = the thread that owns the '>' / the header's newline)
    int64_t* hdr_start;         // offset of '>'
    int64_t* hdr_end;           // offset of the header line's '\n' (or n at EOF)
    int64_t* seq_off;           // rec_cap + 1: sequence bytes kept before record r
    int64_t* gc_prefix;         // rec_cap + 1: G/C among them
    int64_t rec_cap;
    uint8_t* seq_out;           // compacted sequence column (may be null)
    int64_t seq_cap;
};

cudaError_t fasta_scan_launch(const FastaScanArgs& a, int flags, cudaStream_t st);
int64_t fasta_scan_tiles(int64_t begin, int64_t n, int is_final);
int64_t fasta_workspace_payload(int64_t n_tiles);

// ---- writers (writer_ops.cu): columns in HBM -> FASTQ / FASTA file image
struct FormatArgs {
    const int64_t* off[4];    // FASTQ: name, description, sequence, quality; FASTA: id, description, sequence
    const uint8_t* data[4];
    const uint8_t* desc_valid;  // one byte per row, may be null (all valid)
    int64_t n_rows;
    int line_width;           // FASTA
    uint32_t* lens;           // [n_rows] bytes per record in the image
    const int64_t* row_off;   // [n_rows + 1] exclusive scan of lens
    uint8_t* out;
    int64_t out_cap;          // bytes of `out`; a record that does not fit is skipped and counted
    int64_t* long_rows;       // [n_rows] rows above LONG_ROW
    unsigned long long* counters;  // [0] number of long rows, [1] ~(first row whose record exceeds 4 GiB), 0 = none,
                                   // [2] records that did not fit out_cap
};
cudaError_t format_len_launch(const FormatArgs& a, bool fasta, cudaStream_t st);
cudaError_t format_rows_launch(const FormatArgs& a, bool fasta, cudaStream_t st);

}  // namespace exb
