// api.cu -- the device layer of the C ABI (include/exon_b200.h): argument checks,
// workspace carving, kernel launches, synthetic-input generators.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "exon_b200_internal.h"

namespace exb {
// record_ops.cu
cudaError_t exclusive_scan_launch_u32(const uint32_t*, int64_t, int64_t*, TileSlot*, unsigned long long*, cudaStream_t);
cudaError_t exclusive_scan_launch_u32_multi(const uint32_t*, int64_t, int, int64_t, int64_t*, int64_t, TileSlot*, unsigned long long*, cudaStream_t);
cudaError_t exclusive_scan_launch_u8(const uint8_t*, int64_t, int64_t*, TileSlot*, unsigned long long*, cudaStream_t);
int64_t scan_tiles(int64_t n);
cudaError_t select_rows_launch(const uint8_t*, const int64_t*, int64_t, int64_t*, cudaStream_t);
cudaError_t fastq_fields_launch(const uint8_t*, int64_t, int64_t, const void*, bool, const int64_t*, int64_t, uint32_t*, uint8_t*,
                                int64_t*, const ScanResult*, cudaStream_t);
cudaError_t fastq_gather_launch(const uint8_t*, int64_t, int64_t, const void*, bool, const int64_t*, int64_t, int, const uint32_t*,
                                const int64_t*, uint8_t*, cudaStream_t, int, unsigned long long*);
cudaError_t gather_ranges_launch(const uint8_t*, const int64_t*, const uint32_t*, const int64_t*, int64_t, int64_t, uint8_t*, cudaStream_t);
cudaError_t fastq_filter_launch(const uint32_t*, const uint32_t*, const uint32_t*, const int32_t*, int64_t, const exb_predicate*, int,
                                uint8_t*, int64_t*, const void*, cudaStream_t);
cudaError_t fasta_headers_launch(const uint8_t*, const int64_t*, const int64_t*, int64_t, int64_t, uint32_t*, int64_t*, uint8_t*,
                                 unsigned long long*, cudaStream_t);
cudaError_t gc_from_prefix_launch(const int64_t*, const int64_t*, int64_t, float*, cudaStream_t);
cudaError_t gc_from_counts_launch(const uint32_t*, const uint32_t*, int64_t, float*, cudaStream_t);
cudaError_t gc_content_launch(const int64_t*, const uint8_t*, int64_t, float*, cudaStream_t);
cudaError_t seq_map_launch(const uint8_t*, int64_t, int, uint8_t*, unsigned long long*, cudaStream_t);
cudaError_t fastq_final_state_launch(const FastqScanArgs&, cudaStream_t);
cudaError_t translate_launch(const int64_t*, const uint8_t*, int64_t, int64_t, uint8_t*, long long*, cudaStream_t);
cudaError_t quality_decode_launch(const uint8_t*, int64_t, int32_t*, cudaStream_t);
int64_t fastq_split_scratch_bytes(int64_t);
cudaError_t seq_totals_launch(const uint32_t*, const uint32_t*, int64_t, int64_t*, cudaStream_t);
cudaError_t fastq_split_launch(const uint8_t*, int64_t, int64_t, const void*, bool, int64_t, uint32_t, int64_t*, uint8_t*, uint8_t* const*, const int64_t*, void*,
                               const ScanResult*, int, unsigned long long*, cudaStream_t);

static thread_local char g_err[512] = "";
int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return set_err(EXB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// workspace: [0,128) ScanResult | [128,136) ticket | [256, ...) TileSlot[n_tiles]
constexpr int64_t WS_HEADER = 256;
static_assert(sizeof(ScanResult) <= 128, "ScanResult must fit the header");
static_assert(sizeof(ScanResult) == sizeof(exb_scan_result), "ABI mirror");

int64_t tiles_for(int64_t begin, int64_t n, int is_final) {
    int64_t origin = begin & ~(int64_t)15;
    // final chunk: +1 byte of room for the virtual terminator at n
    int64_t t = (n + (is_final ? 1 : 0) - origin + TILE_BYTES - 1) / TILE_BYTES;
    return t > 0 ? t : 1;
}

struct Workspace {
    ScanResult* result;
    unsigned long long* ticket;
    TileSlot* slots;
};
// `payload` bytes of zeroed chain state behind the header
int carve_bytes(void* ws, int64_t ws_bytes, int64_t payload, cudaStream_t st, Workspace* out) {
    int64_t need = WS_HEADER + payload;
    if (!ws || ws_bytes < need) return set_err(EXB_ERR_ARG, "workspace too small: need %lld bytes, have %lld", (long long)need, (long long)ws_bytes);
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)need, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
    out->result = reinterpret_cast<ScanResult*>(ws);
    out->ticket = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws) + 128);
    out->slots = reinterpret_cast<TileSlot*>(reinterpret_cast<uint8_t*>(ws) + WS_HEADER);
    return 0;
}
int carve(void* ws, int64_t ws_bytes, int64_t n_tiles, cudaStream_t st, Workspace* out) {
    return carve_bytes(ws, ws_bytes, n_tiles * (int64_t)sizeof(TileSlot), st, out);
}
// The chain-free scans write every word of their scratch before reading it: same layout, no memset launch.
int carve_scan(void* ws, int64_t ws_bytes, int64_t payload, Workspace* out) {
    const int64_t need = WS_HEADER + payload;
    if (!ws || ws_bytes < need) return set_err(EXB_ERR_ARG, "workspace too small: need %lld bytes, have %lld", (long long)need, (long long)ws_bytes);
    out->result = reinterpret_cast<ScanResult*>(ws);
    out->ticket = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws) + 128);
    out->slots = reinterpret_cast<TileSlot*>(reinterpret_cast<uint8_t*>(ws) + WS_HEADER);
    return 0;
}


// True predecessor state of shard `rank` from the all-gathered result blocks of all shards (one thread: the walk is
// over at most `world` blocks).  Host mirror with the derivation: exon_duckdb_b200/dist.py compose_prev.
__global__ void __launch_bounds__(128) fastq_compose_prev_kernel(const uint8_t* g_blocks, const int64_t* g_ranges, int world, int rank,
                                                                ScanResult* out) {
    // stage the (at most 16) blocks and ranges in shared memory with one round of parallel loads: the walk below is a
    // chain of dependent reads, ~1 us each from global memory and on every rank's critical path once per step
    __shared__ __align__(16) uint64_t s_blocks[16 * 16];
    __shared__ int64_t s_ranges[16 * 3];
    if (world > 16) return;  // the API rejects it
    for (int i = threadIdx.x; i < world * 16; i += blockDim.x) s_blocks[i] = reinterpret_cast<const uint64_t*>(g_blocks)[i];
    for (int i = threadIdx.x; i < world * 3; i += blockDim.x) s_ranges[i] = g_ranges[i];
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint8_t* blocks = reinterpret_cast<const uint8_t*>(s_blocks);
    const int64_t* ranges = s_ranges;
    auto blk = [&](int j) { return reinterpret_cast<const ScanResult*>(blocks + (size_t)j * 128); };
    auto lo = [&](int j) { return ranges[3 * j]; };
    auto hi = [&](int j) { return ranges[3 * j + 1]; };
    auto begin = [&](int j) { return ranges[3 * j + 2]; };
    auto open_start = [&](int j) { return lo(j) + (blk(j)->open_line_start - begin(j)); };  // global offset
    ScanResult r;
    memset(&r, 0, sizeof(r));
    uint64_t lines = 0;
    for (int j = 0; j < rank; j++) lines += blk(j)->total_lines;
    int64_t tail_s = 0, tail_g = 0, start = lo(0);
    for (int j = rank - 1; j >= 0; j--) {  // back to the last shard that saw a newline: the open line starts after it
        tail_s += blk(j)->tail_s;
        tail_g += blk(j)->tail_g;
        if (blk(j)->total_lines > 0) {
            start = open_start(j);
            break;
        }
    }
    uint32_t flags = 0;
    if (start < lo(rank)) {  // the shard that holds the open line's first byte knows what it is
        for (int j = 0; j < rank; j++)
            if (lo(j) <= start && start < hi(j)) {
                flags = blk(j)->pad & 3u;
                break;
            }
    }
    r.total_lines = lines;
    r.open_line_start = begin(rank) + (start - lo(rank));
    r.tail_s = tail_s;
    r.tail_g = tail_g;
    r.pad = flags;
    *out = r;
}
}  // namespace exb

using namespace exb;

extern "C" {
static int64_t fastq_workspace_bytes(int64_t n, int64_t max_lines);

const char* exb_last_error(void) { return g_err; }
const char* exb_version(void) { return "exon-b200 0.1.0 (sm_100a)"; }

int exb_device_available(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n > 0;
}

int64_t exb_scan_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    int64_t cta_tiles = WS_HEADER + (n / TILE_BYTES + 3) * (int64_t)sizeof(TileSlot);  // offset scans
    const int64_t fasta = WS_HEADER + fasta_workspace_payload(fasta_scan_tiles(0, n + 16, 1) + 1);  // FASTA scan
    if (fasta > cta_tiles) cta_tiles = fasta;
    const int64_t fastq = fastq_workspace_bytes(n, n / 24 + 16384);  // FASTQ scan: lines of >= 24 bytes on average
    return cta_tiles > fastq ? cta_tiles : fastq;
}

// FASTQ workspace: [0,256) header (ScanResult | +128 record bump counter)
//                  [256, 256+S) sub-workspace of the offset scan (its own 256-byte header + slots)
//                  tails u64[T] | tile_rec i64[T] | tile_rec2 i64[T] | line_base i64[T+1] | tile_cnt u32[T] | records (8 B each) or fused tiles
struct FastqLayout {
    int64_t n_tiles, scan_ws, off_tails, off_rec, off_rec2, off_base, off_cnt, off_payload, fixed;
};
static FastqLayout fastq_layout(int64_t n_tiles) {
    FastqLayout L;
    L.n_tiles = n_tiles;
    L.scan_ws = WS_HEADER + scan_tiles(n_tiles) * (int64_t)sizeof(TileSlot);
    L.off_tails = WS_HEADER + L.scan_ws;
    L.off_rec = L.off_tails + n_tiles * 8;
    L.off_rec2 = L.off_rec + n_tiles * 8;
    L.off_base = L.off_rec2 + n_tiles * 8;
    L.off_cnt = L.off_base + (n_tiles + 1) * 8;
    L.off_payload = (L.off_cnt + n_tiles * 4 + 15) & ~(int64_t)15;
    L.fixed = L.off_payload;
    return L;
}
static int64_t fastq_workspace_bytes(int64_t n, int64_t max_lines) {
    const FastqLayout L = fastq_layout(fastq_scan_tiles(0, n + 16, 1) + 1);
    const int64_t recs = (max_lines + fastq_record_slack(L.n_tiles)) * 8;
    const int64_t fused = L.n_tiles * (int64_t)sizeof(FusedTile);
    return L.fixed + (recs > fused ? recs : fused) + 64;
}

static int fastq_scan_common(const char* who, const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace,
                             uint64_t max_lines, int flags, void* d_line_end, int64_t line_cap, int wide_offsets, uint32_t* d_seq_len,
                             uint32_t* d_gc, uint32_t* d_qual_len, int32_t* d_qsum, int64_t rec_cap, const exb_predicate* preds, int n_preds,
                             int64_t* d_agg, void* d_workspace, int64_t workspace_bytes, cudaStream_t st, bool resolve_only = false,
                             bool scan_only = false, long long* d_candidates = nullptr) {
    if ((!d_buf && !resolve_only) || begin < 0 || n < begin) return set_err(EXB_ERR_ARG, "%s: bad buffer range", who);
    if (((uintptr_t)d_buf & 15) != 0) return set_err(EXB_ERR_ARG, "%s: d_buf must be 16-byte aligned", who);
    if (d_prev_workspace && (begin & 15) != 0) return set_err(EXB_ERR_ARG, "%s: `begin` of a chained range must be a multiple of 16", who);
    if ((flags & EXB_F_LINES) && !d_line_end) return set_err(EXB_ERR_ARG, "%s: EXB_F_LINES without d_line_end", who);
    if ((flags & EXB_F_LINES) && !wide_offsets && n >= (int64_t)0xFFFFFFFFll)
        return set_err(EXB_ERR_ARG, "%s: 32-bit line offsets need n < 4 GiB", who);
    if (!(flags & EXB_F_FUSED)) {
        if ((flags & EXB_F_SEQ) && (!d_seq_len || !d_gc)) return set_err(EXB_ERR_ARG, "%s: EXB_F_SEQ without outputs", who);
        if ((flags & EXB_F_QUAL) && (!d_qual_len || !d_qsum)) return set_err(EXB_ERR_ARG, "%s: EXB_F_QUAL without outputs", who);
    }
    if (d_prev_workspace == d_workspace) return set_err(EXB_ERR_ARG, "%s: d_prev_workspace must differ from d_workspace", who);
    if (!d_workspace || ((uintptr_t)d_workspace & 15) != 0) return set_err(EXB_ERR_ARG, "%s: d_workspace must be 16-byte aligned", who);
    FastqScanArgs a;
    memset(&a, 0, sizeof(a));
    a.buf = reinterpret_cast<const uint8_t*>(d_buf);
    a.begin = begin;
    a.n = n;
    a.prev = reinterpret_cast<const ScanResult*>(d_prev_workspace);
    a.is_final = is_final ? 1 : 0;
    a.max_lines = max_lines;
    a.n_tiles = fastq_scan_tiles(begin, n, a.is_final);
    const FastqLayout L = fastq_layout(a.n_tiles);
    const int64_t payload = workspace_bytes - L.fixed;
    const bool fused_seq = (flags & EXB_F_FUSED) && (flags & EXB_F_SEQ) && !(flags & EXB_F_QUAL);  // totals flavour: + 32 B per tile
    const int64_t min_payload =
        (flags & EXB_F_FUSED) ? a.n_tiles * (int64_t)(sizeof(FusedTile) + (fused_seq ? 32 : 0)) : fastq_record_slack(a.n_tiles) * 8;
    if (payload < min_payload)
        return set_err(EXB_ERR_ARG, "%s: workspace too small: need at least %lld bytes, have %lld", who, (long long)(L.fixed + min_payload),
                       (long long)workspace_bytes);
    uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
    cudaError_t e;
    if (resolve_only) {  // K1's products stay; only the error mark of the earlier resolution is withdrawn (overflow is K1's)
        e = cudaMemsetAsync(ws + offsetof(ScanResult, err_pos), 0, sizeof(unsigned long long), st);
    } else {
        e = cudaMemsetAsync(ws, 0, (size_t)WS_HEADER, st);  // result block + record bump counter
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace header)");
    a.result = reinterpret_cast<ScanResult*>(ws);
    a.rec_bump = reinterpret_cast<unsigned long long*>(ws + 128);
    a.tails = reinterpret_cast<uint64_t*>(ws + L.off_tails);
    a.tile_rec = reinterpret_cast<int64_t*>(ws + L.off_rec);
    a.tile_rec2 = reinterpret_cast<int64_t*>(ws + L.off_rec2);
    int64_t* line_base = reinterpret_cast<int64_t*>(ws + L.off_base);
    a.line_base = line_base;
    a.tile_cnt = reinterpret_cast<uint32_t*>(ws + L.off_cnt);
    a.records = reinterpret_cast<uint2*>(ws + L.off_payload);
    a.rec_space = payload / 8;
    a.fused_tiles = reinterpret_cast<FusedTile*>(ws + L.off_payload);
    a.fused_fix = reinterpret_cast<unsigned long long*>(ws + L.off_payload + a.n_tiles * (int64_t)sizeof(FusedTile));
    a.n_fused = n_preds;
    for (int i = 0; i < n_preds; i++) a.fused[i] = preds[i];
    a.fused_agg = reinterpret_cast<long long*>(d_agg);
    a.line_end = d_line_end;
    a.line_cap = line_cap;
    a.seq_len = d_seq_len;
    a.gc = d_gc;
    a.qual_len = d_qual_len;
    a.qsum = d_qsum;
    a.rec_cap = rec_cap;
    a.local_records = (flags & EXB_F_LOCAL_RECORDS) ? 1 : 0;
    if (!resolve_only) {
        // K1: every byte once, no inter-tile dependency
        e = fastq_tile_launch(a, flags, st);
        if (e != cudaSuccess) return cuda_fail(e, "fastq_tile launch");
        // global line index of every tile
        Workspace w;
        int rc = carve_scan(ws + WS_HEADER, L.scan_ws, scan_tiles(a.n_tiles) * 8, &w);
        if (rc) return rc;
        e = exclusive_scan_launch_u32(a.tile_cnt, a.n_tiles, line_base, w.slots, w.ticket, st);
        if (e != cudaSuccess) return cuda_fail(e, "line offset scan launch");
    }
    if (d_candidates) {  // K2 for all four phases + the description of the range's first line (single-exchange COUNT)
        e = fastq_candidates_launch(a, d_candidates, st);
        if (e != cudaSuccess) return cuda_fail(e, "fastq_candidates launch");
        return 0;
    }
    if (scan_only) {  // the result block only: K2 runs later through the matching *_resolve call
        e = fastq_final_state_launch(a, st);
        if (e != cudaSuccess) return cuda_fail(e, "fastq_final_state launch");
        return 0;
    }
    // K2: records -> per-line / per-record outputs (or bucket selection, fused)
    e = fastq_emit_launch(a, flags, wide_offsets != 0, st);
    if (e != cudaSuccess) return cuda_fail(e, "fastq_emit launch");
    return 0;
}

int64_t exb_fastq_workspace_bytes(int64_t n, int64_t max_lines) {
    if (n < 0) n = 0;
    if (max_lines < 0) max_lines = 0;
    if (max_lines > n + 1) max_lines = n + 1;
    int64_t generic = WS_HEADER + (n / TILE_BYTES + 3) * (int64_t)sizeof(TileSlot);
    const int64_t fasta = WS_HEADER + fasta_workspace_payload(fasta_scan_tiles(0, n + 16, 1) + 1);
    if (fasta > generic) generic = fasta;
    const int64_t fastq = fastq_workspace_bytes(n, max_lines);
    return generic > fastq ? generic : fastq;
}

int exb_fastq_scan(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, uint64_t max_lines,
                   int flags, void* d_line_end, int64_t line_cap, int wide_offsets, uint32_t* d_seq_len, uint32_t* d_gc,
                   uint32_t* d_qual_len, int32_t* d_qsum, int64_t rec_cap, void* d_workspace, int64_t workspace_bytes, void* stream) {
    return fastq_scan_common("exb_fastq_scan", d_buf, begin, n, is_final, d_prev_workspace, max_lines, flags & (7 | EXB_F_LOCAL_RECORDS), d_line_end, line_cap,
                             wide_offsets, d_seq_len, d_gc, d_qual_len, d_qsum, rec_cap, nullptr, 0, nullptr, d_workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

int exb_fastq_scan_filter(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace,
                          const exb_predicate* preds, int n_preds, int64_t* d_agg, int accumulate, void* d_workspace,
                          int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_preds < 0 || n_preds > EXB_MAX_PREDICATES) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter: at most %d predicates", EXB_MAX_PREDICATES);
    for (int i = 0; i < n_preds; i++) {
        if (preds[i].field != EXB_P_MEAN_QUALITY && preds[i].field != EXB_P_QUAL_LEN)
            return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter: predicate %d is not on the quality line (use exb_fastq_scan + exb_fastq_filter)", i);
        if (preds[i].op < 0 || preds[i].op > 5) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter: bad operator in predicate %d", i);
    }
    if (!d_agg) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter: d_agg is required");
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(d_agg, 0, 8 * sizeof(int64_t), st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(agg)");
    }
    return fastq_scan_common("exb_fastq_scan_filter", d_buf, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_QUAL, nullptr, 0,
                             0, nullptr, nullptr, nullptr, nullptr, 0, preds, n_preds, d_agg, d_workspace, workspace_bytes, st);
}

// ---- fused TOTALS flavour (BASELINE C5): SUM(length(sequence)), SUM(#GC), AVG(gc_content(sequence)) in the byte pass
int exb_fastq_scan_totals(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, int64_t* d_agg,
                          int accumulate, void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_agg) return set_err(EXB_ERR_ARG, "exb_fastq_scan_totals: d_agg is null");
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(d_agg, 0, 8 * sizeof(int64_t), st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(agg)");
    }
    return fastq_scan_common("exb_fastq_scan_totals", d_buf, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_SEQ, nullptr, 0, 0,
                             nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, d_agg, d_workspace, workspace_bytes, st);
}
int exb_fastq_scan_totals_begin(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, void* d_workspace,
                                int64_t workspace_bytes, void* stream) {
    return fastq_scan_common("exb_fastq_scan_totals_begin", d_buf, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_SEQ, nullptr, 0, 0,
                             nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream, false, true);
}
int exb_fastq_scan_totals_resolve(int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, int64_t* d_agg, int accumulate,
                                  void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_agg) return set_err(EXB_ERR_ARG, "exb_fastq_scan_totals_resolve: d_agg is null");
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(d_agg, 0, 8 * sizeof(int64_t), st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(agg)");
    }
    return fastq_scan_common("exb_fastq_scan_totals_resolve", nullptr, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_SEQ, nullptr, 0,
                             0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, d_agg, d_workspace, workspace_bytes, st, true);
}

int exb_fastq_scan_filter_begin(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace,
                                const exb_predicate* preds, int n_preds, void* d_workspace, int64_t workspace_bytes, void* stream) {
    if (n_preds < 0 || n_preds > EXB_MAX_PREDICATES || (n_preds > 0 && !preds)) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_begin: bad predicates");
    for (int i = 0; i < n_preds; i++) {
        if (preds[i].field != EXB_P_MEAN_QUALITY && preds[i].field != EXB_P_QUAL_LEN)
            return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_begin: predicate %d is not on the quality line", i);
        if (preds[i].op < 0 || preds[i].op > 5) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_begin: bad operator in predicate %d", i);
    }
    // the byte pass judges every line under all four phase hypotheses, so it needs the predicates; K2 only picks buckets
    return fastq_scan_common("exb_fastq_scan_filter_begin", d_buf, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_QUAL, nullptr, 0,
                             0, nullptr, nullptr, nullptr, nullptr, 0, preds, n_preds, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream,
                             false, true);
}

int exb_fastq_scan_filter_candidates(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace,
                                     const exb_predicate* preds, int n_preds, void* d_record, void* d_workspace, int64_t workspace_bytes,
                                     void* stream) {
    if (n_preds < 0 || n_preds > EXB_MAX_PREDICATES || (n_preds > 0 && !preds) || !d_record)
        return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_candidates: bad arguments");
    for (int i = 0; i < n_preds; i++) {
        if (preds[i].field != EXB_P_MEAN_QUALITY && preds[i].field != EXB_P_QUAL_LEN)
            return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_candidates: predicate %d is not on the quality line", i);
        if (preds[i].op < 0 || preds[i].op > 5) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_candidates: bad operator in predicate %d", i);
    }
    return fastq_scan_common("exb_fastq_scan_filter_candidates", d_buf, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_QUAL, nullptr, 0,
                             0, nullptr, nullptr, nullptr, nullptr, 0, preds, n_preds, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream,
                             false, true, reinterpret_cast<long long*>(d_record));
}

int exb_fastq_combine_records(const void* d_records, const int64_t* d_ranges, int world, const exb_predicate* preds, int n_preds,
                              int64_t* d_total, void* stream) {
    if (!d_records || !d_ranges || !d_total || world < 1 || world > 16 || n_preds < 0 || n_preds > EXB_MAX_PREDICATES)
        return set_err(EXB_ERR_ARG, "exb_fastq_combine_records: bad arguments (1 <= world <= 16)");
    cudaError_t e = fastq_combine_records_launch(reinterpret_cast<const long long*>(d_records), d_ranges, world, preds, n_preds,
                                                 reinterpret_cast<long long*>(d_total), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "fastq_combine_records launch");
    return 0;
}

int exb_fastq_scan_begin(const void* d_buf, int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, int flags, void* d_workspace,
                         int64_t workspace_bytes, void* stream) {
    // K1 writes nothing per record: the outputs are only validated by the resolve call that runs K2
    static uint32_t dummy;
    uint32_t* d = &dummy;  // non-null placeholder for the argument check; K1 never dereferences the output arrays
    return fastq_scan_common("exb_fastq_scan_begin", d_buf, begin, n, is_final, d_prev_workspace, ~0ull, flags & 7, (flags & EXB_F_LINES) ? d : nullptr, 0,
                             1, d, d, d, reinterpret_cast<int32_t*>(d), 0, nullptr, 0, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream, false, true);
}

int exb_fastq_scan_resolve(int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, uint64_t max_lines, int flags,
                           void* d_line_end, int64_t line_cap, int wide_offsets, uint32_t* d_seq_len, uint32_t* d_gc, uint32_t* d_qual_len,
                           int32_t* d_qsum, int64_t rec_cap, void* d_workspace, int64_t workspace_bytes, void* stream) {
    return fastq_scan_common("exb_fastq_scan_resolve", nullptr, begin, n, is_final, d_prev_workspace, max_lines, flags & (7 | EXB_F_LOCAL_RECORDS), d_line_end, line_cap,
                             wide_offsets, d_seq_len, d_gc, d_qual_len, d_qsum, rec_cap, nullptr, 0, nullptr, d_workspace, workspace_bytes,
                             (cudaStream_t)stream, true);
}

int exb_fastq_scan_filter_resolve(int64_t begin, int64_t n, int is_final, const void* d_prev_workspace, const exb_predicate* preds,
                                  int n_preds, int64_t* d_agg, int accumulate, void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_preds < 0 || n_preds > EXB_MAX_PREDICATES || !d_agg) return set_err(EXB_ERR_ARG, "exb_fastq_scan_filter_resolve: bad arguments");
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(d_agg, 0, 8 * sizeof(int64_t), st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(agg)");
    }
    return fastq_scan_common("exb_fastq_scan_filter_resolve", nullptr, begin, n, is_final, d_prev_workspace, ~0ull, EXB_F_FUSED | EXB_F_QUAL,
                             nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, preds, n_preds, d_agg, d_workspace, workspace_bytes, st, true);
}

int exb_fastq_compose_prev(const void* d_blocks, const int64_t* d_ranges, int world, int rank, void* d_prev_out, void* stream) {
    if (!d_blocks || !d_ranges || !d_prev_out || world < 1 || world > 16 || rank < 0 || rank >= world)
        return set_err(EXB_ERR_ARG, "exb_fastq_compose_prev: bad arguments (1 <= world <= 16, 0 <= rank < world)");
    fastq_compose_prev_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(d_blocks), d_ranges, world, rank,
                                                                  reinterpret_cast<ScanResult*>(d_prev_out));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fastq_compose_prev launch");
    return 0;
}

int exb_scan_result_store(void* d_dst, const exb_scan_result* src, void* stream) {
    if (!d_dst || !src) return set_err(EXB_ERR_ARG, "exb_scan_result_store: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    exb_scan_result tmp = *src;
    tmp.err_pos = ~tmp.err_pos;  // device representation (see exb_scan_result_fetch)
    // pageable source: the runtime stages the bytes before the call returns, so `tmp` may go out of scope
    cudaError_t e = cudaMemcpyAsync(d_dst, &tmp, sizeof(tmp), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(result store)");
    return 0;
}

int exb_scan_result_fetch(const void* d_workspace, exb_scan_result* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(out, d_workspace, sizeof(exb_scan_result), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(result)");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    out->err_pos = ~out->err_pos;  // the device keeps it inverted (see ScanResult)
    return 0;
}

int exb_fastq_filter(const uint32_t* d_seq_len, const uint32_t* d_gc, const uint32_t* d_qual_len, const int32_t* d_qsum,
                     int64_t n_records, const exb_predicate* preds, int n_preds, uint8_t* d_pass, int64_t* d_agg,
                     const void* d_scan_workspace, void* stream) {
    if (n_preds < 0 || n_preds > EXB_MAX_PREDICATES) return set_err(EXB_ERR_ARG, "exb_fastq_filter: at most %d predicates", EXB_MAX_PREDICATES);
    for (int i = 0; i < n_preds; i++) {
        int f = preds[i].field;
        if (f < 0 || f > 3 || preds[i].op < 0 || preds[i].op > 5) return set_err(EXB_ERR_ARG, "exb_fastq_filter: bad predicate %d", i);
        if (f == EXB_P_MEAN_QUALITY && (!d_qual_len || !d_qsum)) return set_err(EXB_ERR_ARG, "mean_quality predicate needs qual_len/qsum");
        if (f == EXB_P_GC_CONTENT && (!d_seq_len || !d_gc)) return set_err(EXB_ERR_ARG, "gc_content predicate needs seq_len/gc");
        if (f == EXB_P_SEQ_LEN && !d_seq_len) return set_err(EXB_ERR_ARG, "length(sequence) predicate needs seq_len");
        if (f == EXB_P_QUAL_LEN && !d_qual_len) return set_err(EXB_ERR_ARG, "length(quality_scores) predicate needs qual_len");
    }
    if (!d_agg) return set_err(EXB_ERR_ARG, "exb_fastq_filter: d_agg is required");
    cudaError_t e = fastq_filter_launch(d_seq_len, d_gc, d_qual_len, d_qsum, n_records, preds, n_preds, d_pass, d_agg, d_scan_workspace,
                                        (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "fastq_filter launch");
    return 0;
}

int exb_fastq_fields(const void* d_buf, int64_t begin, int64_t n, const void* d_line_end, int wide_offsets, const int64_t* d_sel,
                     int64_t n_rows, uint32_t* d_lens, uint8_t* d_desc_valid, int64_t* d_starts, const void* d_scan_workspace, void* stream) {
    cudaError_t e = fastq_fields_launch(reinterpret_cast<const uint8_t*>(d_buf), begin, n, d_line_end, wide_offsets != 0, d_sel, n_rows,
                                        d_lens, d_desc_valid, d_starts, reinterpret_cast<const ScanResult*>(d_scan_workspace),
                                        (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "fastq_fields launch");
    return 0;
}

int exb_exclusive_scan_u32(const uint32_t* d_in, int64_t n, int64_t* d_out, void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w;
    int rc = carve_scan(d_workspace, workspace_bytes, scan_tiles(n) * 8, &w);
    if (rc) return rc;
    cudaError_t e = exclusive_scan_launch_u32(d_in, n, d_out, w.slots, w.ticket, st);
    if (e != cudaSuccess) return cuda_fail(e, "exclusive_scan launch");
    return 0;
}

int exb_exclusive_scan_u32_multi(const uint32_t* d_in, int64_t n, int cols, int64_t in_stride, int64_t* d_out, int64_t out_stride,
                                 void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (cols < 1 || cols > 8) return set_err(EXB_ERR_ARG, "exb_exclusive_scan_u32_multi: cols must be 1..8");
    Workspace w;  // chains are one 8-byte word per tile and column, packed column after column; tickets sit in the header
    int rc = carve_scan(d_workspace, workspace_bytes, (int64_t)cols * scan_tiles(n) * 8 + 64, &w);
    if (rc) return rc;
    cudaError_t e = exclusive_scan_launch_u32_multi(d_in, n, cols, in_stride, d_out, out_stride, w.slots, w.ticket, st);
    if (e != cudaSuccess) return cuda_fail(e, "exclusive_scan launch");
    return 0;
}

int exb_select_rows(const uint8_t* d_pass, int64_t n, int64_t* d_offsets, int64_t* d_sel, void* d_workspace, int64_t workspace_bytes,
                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w;
    int rc = carve_scan(d_workspace, workspace_bytes, scan_tiles(n) * 8, &w);
    if (rc) return rc;
    cudaError_t e = exclusive_scan_launch_u8(d_pass, n, d_offsets, w.slots, w.ticket, st);
    if (e != cudaSuccess) return cuda_fail(e, "exclusive_scan launch");
    e = select_rows_launch(d_pass, d_offsets, n, d_sel, st);
    if (e != cudaSuccess) return cuda_fail(e, "select_rows launch");
    return 0;
}

int exb_fastq_gather(const void* d_buf, int64_t begin, int64_t n, const void* d_line_end, int wide_offsets, const int64_t* d_sel,
                     int64_t n_rows, int col, const uint32_t* d_lens, const int64_t* d_off, uint8_t* d_out, void* stream) {
    if (col < 0 || col > 3) return set_err(EXB_ERR_ARG, "exb_fastq_gather: col must be 0..3");
    cudaError_t e = fastq_gather_launch(reinterpret_cast<const uint8_t*>(d_buf), begin, n, d_line_end, wide_offsets != 0, d_sel, n_rows,
                                        col, d_lens, d_off, d_out, (cudaStream_t)stream, 0, nullptr);
    if (e != cudaSuccess) return cuda_fail(e, "fastq_gather launch");
    return 0;
}

int64_t exb_fastq_split_scratch_bytes(int64_t n_rows) { return fastq_split_scratch_bytes(n_rows); }
int exb_fastq_split(const void* d_buf, int64_t begin, int64_t n, const void* d_line_end, int wide_offsets, int64_t n_rows, uint32_t column_mask,
                    int64_t* d_off, uint8_t* d_desc_valid, uint8_t* const* d_out, const int64_t* cap, void* d_scratch, const void* d_scan_workspace,
                    int map_mode, uint64_t* d_bad, void* stream) {
    if (!d_buf || !d_line_end || !d_off || !d_desc_valid || !d_out || !cap || !d_scratch || n_rows < 0)
        return set_err(EXB_ERR_ARG, "exb_fastq_split: null argument");
    uint8_t* out[4];
    int64_t caps[4];
    for (int c = 0; c < 4; c++) {
        out[c] = d_out[c];
        caps[c] = cap[c];
        if (((column_mask >> c) & 1u) && (!out[c] || ((uintptr_t)out[c] & 15))) return set_err(EXB_ERR_ARG, "exb_fastq_split: column %d needs a 16-byte aligned output", c);
    }
    if (((uintptr_t)d_buf & 15) != 0) return set_err(EXB_ERR_ARG, "exb_fastq_split: d_buf must be 16-byte aligned");
    if (map_mode >= 0 && (!d_bad || map_mode > EXB_MAP_REVERSE_TRANSCRIBE)) return set_err(EXB_ERR_ARG, "exb_fastq_split: bad map arguments");
    cudaError_t e = fastq_split_launch(reinterpret_cast<const uint8_t*>(d_buf), begin, n, d_line_end, wide_offsets != 0, n_rows, column_mask, d_off,
                                       d_desc_valid, out, caps, d_scratch, reinterpret_cast<const ScanResult*>(d_scan_workspace), map_mode,
                                       map_mode >= 0 ? reinterpret_cast<unsigned long long*>(d_bad) : nullptr, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "fastq_split launch");
    return 0;
}

int exb_fastq_gather_map(const void* d_buf, int64_t begin, int64_t n, const void* d_line_end, int wide_offsets, const int64_t* d_sel,
                         int64_t n_rows, int col, const uint32_t* d_lens, const int64_t* d_off, int mode, uint8_t* d_out, uint64_t* d_bad,
                         void* stream) {
    if (col < 2 || col > 3) return set_err(EXB_ERR_ARG, "exb_fastq_gather_map: col must be 2 (sequence) or 3 (quality_scores)");
    if (mode < EXB_MAP_REVERSE_COMPLEMENT || mode > EXB_MAP_REVERSE_TRANSCRIBE || !d_bad)
        return set_err(EXB_ERR_ARG, "exb_fastq_gather_map: mode must be one of EXB_MAP_*, d_bad non-null");
    cudaError_t e = fastq_gather_launch(reinterpret_cast<const uint8_t*>(d_buf), begin, n, d_line_end, wide_offsets != 0, d_sel, n_rows,
                                        col, d_lens, d_off, d_out, (cudaStream_t)stream, mode, reinterpret_cast<unsigned long long*>(d_bad));
    if (e != cudaSuccess) return cuda_fail(e, "fastq_gather_map launch");
    return 0;
}

int exb_fasta_scan(const void* d_buf, int64_t begin, int64_t n, int is_final, int64_t halo_n, const void* d_prev_workspace,
                   int64_t* d_hdr_start, int64_t* d_hdr_end, int64_t* d_seq_off, int64_t* d_gc_prefix, int64_t rec_cap,
                   uint8_t* d_seq_out, int64_t seq_cap, void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_buf || begin < 0 || n < begin) return set_err(EXB_ERR_ARG, "exb_fasta_scan: bad buffer range");
    if (((uintptr_t)d_buf & 15) != 0) return set_err(EXB_ERR_ARG, "exb_fasta_scan: d_buf must be 16-byte aligned");
    if (d_seq_out && ((uintptr_t)d_seq_out & 15) != 0) return set_err(EXB_ERR_ARG, "exb_fasta_scan: d_seq_out must be 16-byte aligned");
    if (!d_hdr_start || !d_hdr_end || !d_seq_off || !d_gc_prefix) return set_err(EXB_ERR_ARG, "exb_fasta_scan: per-record outputs are required");
    FastaScanArgs a;
    memset(&a, 0, sizeof(a));
    a.buf = reinterpret_cast<const uint8_t*>(d_buf);
    a.begin = begin;
    a.n = n;
    a.prev = reinterpret_cast<const ScanResult*>(d_prev_workspace);
    a.is_final = is_final ? 1 : 0;
    a.halo_n = halo_n > n ? halo_n : n;
    a.n_tiles = fasta_scan_tiles(begin, n, a.is_final);
    if (d_prev_workspace == d_workspace) return set_err(EXB_ERR_ARG, "exb_fasta_scan: d_prev_workspace must differ from d_workspace");
    if (d_prev_workspace && (begin & 15) != 0) return set_err(EXB_ERR_ARG, "exb_fasta_scan: `begin` of a chained range must be a multiple of 16");
    if (!d_workspace || ((uintptr_t)d_workspace & 15) != 0) return set_err(EXB_ERR_ARG, "exb_fasta_scan: d_workspace must be 16-byte aligned");
    const int64_t need = WS_HEADER + fasta_workspace_payload(a.n_tiles);
    if (workspace_bytes < need)
        return set_err(EXB_ERR_ARG, "exb_fasta_scan: workspace too small: need %lld bytes, have %lld", (long long)need, (long long)workspace_bytes);
    cudaError_t e0 = cudaMemsetAsync(d_workspace, 0, (size_t)WS_HEADER, st);  // result block (err_pos / overflow accumulate)
    if (e0 != cudaSuccess) return cuda_fail(e0, "cudaMemsetAsync(workspace header)");
    a.result = reinterpret_cast<ScanResult*>(d_workspace);
    a.payload = reinterpret_cast<uint8_t*>(d_workspace) + WS_HEADER;
    a.hdr_start = d_hdr_start;
    a.hdr_end = d_hdr_end;
    a.seq_off = d_seq_off;
    a.gc_prefix = d_gc_prefix;
    a.rec_cap = rec_cap;
    a.seq_out = d_seq_out;
    a.seq_cap = seq_cap;
    cudaError_t e = fasta_scan_launch(a, 0, st);
    if (e != cudaSuccess) return cuda_fail(e, "fasta_scan launch");
    return 0;
}

int exb_fasta_headers(const void* d_buf, int64_t n, const int64_t* d_hdr_start, const int64_t* d_hdr_end, int64_t n_rows,
                      uint32_t* d_lens, int64_t* d_desc_start, uint8_t* d_desc_valid, uint64_t* d_err_pos, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_err_pos, 0xFF, 8, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    e = fasta_headers_launch(reinterpret_cast<const uint8_t*>(d_buf), d_hdr_start, d_hdr_end, n_rows, n, d_lens, d_desc_start,
                             d_desc_valid, reinterpret_cast<unsigned long long*>(d_err_pos), st);
    if (e != cudaSuccess) return cuda_fail(e, "fasta_headers launch");
    return 0;
}

int exb_gather_ranges(const void* d_buf, const int64_t* d_start, const uint32_t* d_len, const int64_t* d_off, int64_t n_rows,
                      uint8_t* d_out, int64_t out_bytes, void* stream) {
    if (n_rows < 0 || out_bytes < 0 || (n_rows > 0 && (!d_buf || !d_start || !d_off || !d_out)))
        return set_err(EXB_ERR_ARG, "exb_gather_ranges: bad arguments");
    cudaError_t e = gather_ranges_launch(reinterpret_cast<const uint8_t*>(d_buf), d_start, d_len, d_off, n_rows, out_bytes, d_out,
                                         (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gather_ranges launch");
    return 0;
}

int exb_gc_from_prefix(const int64_t* d_seq_off, const int64_t* d_gc_prefix, int64_t n_rows, float* d_out, void* stream) {
    cudaError_t e = gc_from_prefix_launch(d_seq_off, d_gc_prefix, n_rows, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gc_from_prefix launch");
    return 0;
}
int exb_gc_from_counts(const uint32_t* d_seq_len, const uint32_t* d_gc, int64_t n_rows, float* d_out, void* stream) {
    cudaError_t e = gc_from_counts_launch(d_seq_len, d_gc, n_rows, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gc_from_counts launch");
    return 0;
}
int exb_fastq_seq_totals(const uint32_t* d_seq_len, const uint32_t* d_gc, int64_t n_records, int64_t* d_totals, void* stream) {
    if (!d_seq_len || !d_gc || !d_totals || n_records < 0) return set_err(EXB_ERR_ARG, "exb_fastq_seq_totals: bad arguments");
    cudaError_t e = seq_totals_launch(d_seq_len, d_gc, n_records, d_totals, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "seq_totals launch");
    return 0;
}

int exb_gc_content(const int64_t* d_off, const uint8_t* d_data, int64_t n_rows, float* d_out, void* stream) {
    cudaError_t e = gc_content_launch(d_off, d_data, n_rows, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gc_content launch");
    return 0;
}
int exb_seq_map(const uint8_t* d_in, int64_t n_bytes, int mode, uint8_t* d_out, uint64_t* d_bad_pos, void* stream) {
    if (mode < EXB_MAP_REVERSE_COMPLEMENT || mode > EXB_MAP_REVERSE_TRANSCRIBE) return set_err(EXB_ERR_ARG, "exb_seq_map: bad mode");
    cudaError_t e = seq_map_launch(d_in, n_bytes, mode, d_out, reinterpret_cast<unsigned long long*>(d_bad_pos), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "seq_map launch");
    return 0;
}
static int exb_translate_bound(const int64_t* d_off, const uint8_t* d_data, int64_t n_rows, int64_t n_bytes_bound, uint8_t* d_out, int64_t* d_status,
                        void* stream) {
    if (n_rows < 0 || !d_status || (n_rows > 0 && (!d_off || !d_out))) return set_err(EXB_ERR_ARG, "exb_translate: bad arguments");
    cudaError_t e = translate_launch(d_off, d_data, n_rows, n_bytes_bound, d_out, reinterpret_cast<long long*>(d_status), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "translate launch");
    return 0;
}
int exb_translate(const int64_t* d_off, const uint8_t* d_data, int64_t n_rows, uint8_t* d_out, int64_t* d_status, void* stream) {
    // the column's size is on the device: size the grid for the machine (the kernel is grid-strided)
    return exb_translate_bound(d_off, d_data, n_rows, (int64_t)148 * 16 * 256 * 3, d_out, d_status, stream);
}

int exb_quality_decode(const uint8_t* d_in, int64_t n_bytes, int32_t* d_out, void* stream) {
    cudaError_t e = quality_decode_launch(d_in, n_bytes, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "quality_decode launch");
    return 0;
}

}  // extern "C"
