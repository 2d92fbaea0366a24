// record_ops.cu -- per-record kernels that consume the scan outputs:
// predicate / aggregate evaluation, field extents, offset scans, gathers into
// Arrow-style column buffers, and the sequence scalar functions.
//
// Reference semantics restated here:
//   gc_content                 exon/src/exon/sequence_functions/module.cpp:131-158
//   reverse_complement         module.cpp:30-69   (A->C T->G C->A G->T, no reversal)
//   complement                 module.cpp:81-121
//   quality_score_string_to_list   exon/src/exon/fastq_functions/module.cpp:32-50
//   name/description split     noodles-fastq 0.8 read_record / noodles-fasta 0.27
//                              Definition::from_str (not in the reference tree)
#include "common.cuh"
#include "exon_b200_internal.h"
#include "tma_tile.cuh"
#include "x87div.h"

namespace exb {

// ================================================================= filter
struct FilterArgs {
    const uint32_t *seq_len, *gc, *qual_len;
    const int32_t* qsum;
    int64_t n;
    exb_predicate preds[EXB_MAX_PREDICATES];
    int n_preds;
    uint8_t* pass;
    long long* agg;
    const ScanResult* scan;  // optional: take the record count from a scan's result block (no host round trip)
};

__device__ __forceinline__ float gc_fraction(uint32_t gc, uint32_t len) {
    // (float)gc_count / (float)size with both operands C ints; '' -> 0 (module.cpp:142-156)
    return len == 0 ? 0.0f : __fdiv_rn(__uint2float_rn(gc), __uint2float_rn(len));
}

__global__ void __launch_bounds__(256) fastq_filter_kernel(FilterArgs a) {
    long long v[5] = {0, 0, 0, 0, 0};
    int64_t n = a.n;
    if (a.scan) {
        const int64_t have = (int64_t)(a.scan->total_lines >> 2);
        n = have < n ? have : n;
    }
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t sl = a.seq_len ? a.seq_len[r] : 0, g = a.gc ? a.gc[r] : 0;
        const uint32_t ql = a.qual_len ? a.qual_len[r] : 0;
        const int32_t qs = a.qsum ? a.qsum[r] : 0;
        bool ok = true;
        for (int i = 0; i < a.n_preds; i++) {
            const exb_predicate p = a.preds[i];
            bool c;
            switch (p.field) {
            case EXB_P_MEAN_QUALITY: c = exb_mean_cmp((int64_t)qs, ql, p.op, p.value); break;
            case EXB_P_GC_CONTENT: c = exb_cmp((double)gc_fraction(g, sl), p.op, p.value); break;
            case EXB_P_SEQ_LEN: c = exb_cmp((double)sl, p.op, p.value); break;
            default: c = exb_cmp((double)ql, p.op, p.value); break;
            }
            ok = ok && c;
        }
        if (a.pass) a.pass[r] = ok ? 1 : 0;
        if (ok) {
            v[0] += 1;
            v[1] += sl;
            v[2] += g;
            v[3] += qs;
            v[4] += ql;
        }
    }
    __shared__ long long s_acc[5];
    if (threadIdx.x < 5) s_acc[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 5; i++) {
        long long x = v[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0 && x != 0) atomicAdd((unsigned long long*)&s_acc[i], (unsigned long long)x);
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_acc[threadIdx.x] != 0)
        atomicAdd((unsigned long long*)&a.agg[threadIdx.x], (unsigned long long)s_acc[threadIdx.x]);
}

// ================================================================= exclusive scan (u32 / u8 -> i64)
// Chain-free: reduce (one sum per tile of 4096 items) -> spine (one block per column scans the tile sums in place)
// -> downsweep (local scan + tile base).  The first version was a single-pass decoupled look-back; with 16 KiB tiles a
// hop costs ~140 ns and the chain serialises (58 us for the 1.7 M per-tile line counts of a 7 GB FASTQ, 150 us for the
// four 4 M-row offset columns of a table).  Three short launches read the input twice but never wait on a neighbour.
// `cols` independent columns share the launches (blockIdx.y): column c reads in + c * in_stride, writes
// out + c * out_stride and keeps its tile sums at sums + c * n_tiles.
constexpr int SCAN_VEC = 4, SCAN_ROUNDS = 4;
constexpr int SCAN_ROUND_ITEMS = BLOCK_THREADS * SCAN_VEC;   // 1024 items per round: a thread owns 4 consecutive ones
constexpr int SCAN_TILE = SCAN_ROUND_ITEMS * SCAN_ROUNDS;   // 4096 items per tile

// four consecutive items as one 16-byte (u32) / 4-byte (u8) load when the column is aligned for it, so a warp reads
// 512 (128) contiguous bytes per instruction; the first version read 16 items per thread with 64-byte strides
template <typename T>
__device__ __forceinline__ uint32_t scan_load4(const T* __restrict__ in, int64_t idx, int64_t n, bool vec_ok, uint32_t (&x)[SCAN_VEC]) {
    if (vec_ok && idx + SCAN_VEC <= n) {
        if (sizeof(T) == 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(in + idx);
            x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
        } else {
            const uchar4 v = *reinterpret_cast<const uchar4*>(in + idx);
            x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_VEC; i++) x[i] = (idx + i < n) ? (uint32_t)in[idx + i] : 0u;
    }
    return x[0] + x[1] + x[2] + x[3];
}

template <typename T>
__global__ void __launch_bounds__(BLOCK_THREADS) scan_reduce_kernel(const T* __restrict__ in, int64_t n, uint64_t* __restrict__ sums, int64_t n_tiles,
                                                                   int64_t in_stride) {
    __shared__ uint64_t s_warp[WARPS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    in += (int64_t)blockIdx.y * in_stride;
    sums += (int64_t)blockIdx.y * n_tiles;
    const bool vec_ok = ((uintptr_t)in & (SCAN_VEC * sizeof(T) - 1)) == 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t loc = 0;
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            uint32_t x[SCAN_VEC];
            loc += scan_load4(in, tile * SCAN_TILE + r * SCAN_ROUND_ITEMS + t * SCAN_VEC, n, vec_ok, x);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, d);
        __syncthreads();
        if (lane == 0) s_warp[warp] = loc;
        __syncthreads();
        if (t == 0) {
            uint64_t tot = 0;
            for (int w = 0; w < WARPS; w++) tot += s_warp[w];
            sums[tile] = tot;
        }
    }
}

// one block per column: tile sums -> exclusive tile bases, in place; the column total goes to out[n]
__global__ void __launch_bounds__(1024) scan_spine_kernel(uint64_t* __restrict__ sums, int64_t n_tiles, int64_t* __restrict__ out, int64_t n,
                                                         int64_t out_stride) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_run;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    sums += (int64_t)blockIdx.x * n_tiles;
    out += (int64_t)blockIdx.x * out_stride;
    if (t == 0) s_run = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + t;
        const uint64_t v = i < n_tiles ? sums[i] : 0ull;
        const uint64_t incl = warp_incl_scan_u64(v);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint64_t woff = 0, tot = 0;
        for (int w = 0; w < 32; w++) {
            if (w < warp) woff += s_warp[w];
            tot += s_warp[w];
        }
        const uint64_t run = s_run;
        if (i < n_tiles) sums[i] = run + woff + incl - v;
        __syncthreads();
        if (t == 0) s_run = run + tot;
        __syncthreads();
    }
    if (t == 0) out[n] = (int64_t)s_run;
}

template <typename T>
__global__ void __launch_bounds__(BLOCK_THREADS) scan_down_kernel(const T* __restrict__ in, int64_t n, int64_t* __restrict__ out,
                                                                 const uint64_t* __restrict__ sums, int64_t n_tiles, int64_t in_stride,
                                                                 int64_t out_stride) {
    __shared__ uint64_t s_warp[2][WARPS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    in += (int64_t)blockIdx.y * in_stride;
    out += (int64_t)blockIdx.y * out_stride;
    sums += (int64_t)blockIdx.y * n_tiles;
    const bool vec_ok = ((uintptr_t)in & (SCAN_VEC * sizeof(T) - 1)) == 0;
    const bool out16 = ((uintptr_t)out & 15) == 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t carry = sums[tile];
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            const int64_t idx = tile * SCAN_TILE + r * SCAN_ROUND_ITEMS + t * SCAN_VEC;
            uint32_t x[SCAN_VEC];
            const uint64_t loc = scan_load4(in, idx, n, vec_ok, x);
            const uint64_t incl = warp_incl_scan_u64(loc);
            if (lane == 31) s_warp[r & 1][warp] = incl;  // double-buffered: one barrier per round
            __syncthreads();
            uint64_t woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const uint64_t v = s_warp[r & 1][w];
                if (w < warp) woff += v;
                tot += v;
            }
            const int64_t o0 = (int64_t)(carry + woff + incl - loc), o1 = o0 + x[0], o2 = o1 + x[1], o3 = o2 + x[2];
            if (idx + SCAN_VEC <= n && out16) {
                reinterpret_cast<longlong2*>(out + idx)[0] = make_longlong2(o0, o1);
                reinterpret_cast<longlong2*>(out + idx)[1] = make_longlong2(o2, o3);
            } else {
                if (idx < n) out[idx] = o0;
                if (idx + 1 < n) out[idx + 1] = o1;
                if (idx + 2 < n) out[idx + 2] = o2;
                if (idx + 3 < n) out[idx + 3] = o3;
            }
            carry += tot;
        }
        __syncthreads();  // the next tile reuses s_warp[0]
    }
}

int64_t scan_tiles(int64_t n) {
    int64_t t = (n + SCAN_TILE - 1) / SCAN_TILE;
    return t ? t : 1;
}

template <typename T>
static cudaError_t exclusive_scan_launch_t(const T* in, int64_t n, int cols, int64_t in_stride, int64_t* out, int64_t out_stride, TileSlot* slots,
                                           cudaStream_t st) {
    const int64_t n_tiles = scan_tiles(n);
    uint64_t* sums = reinterpret_cast<uint64_t*>(slots);  // cols * n_tiles words of the caller's workspace
    const unsigned gx = (unsigned)(n_tiles < 148 * 16 ? n_tiles : 148 * 16);
    scan_reduce_kernel<T><<<dim3(gx, (unsigned)cols), BLOCK_THREADS, 0, st>>>(in, n, sums, n_tiles, in_stride);
    scan_spine_kernel<<<(unsigned)cols, 1024, 0, st>>>(sums, n_tiles, out, n, out_stride);
    scan_down_kernel<T><<<dim3(gx, (unsigned)cols), BLOCK_THREADS, 0, st>>>(in, n, out, sums, n_tiles, in_stride, out_stride);
    return cudaGetLastError();
}
cudaError_t exclusive_scan_launch_u32(const uint32_t* in, int64_t n, int64_t* out, TileSlot* slots, unsigned long long* ticket,
                                      cudaStream_t st) {
    (void)ticket;
    return exclusive_scan_launch_t<uint32_t>(in, n, 1, 0, out, 0, slots, st);
}
cudaError_t exclusive_scan_launch_u32_multi(const uint32_t* in, int64_t n, int cols, int64_t in_stride, int64_t* out, int64_t out_stride,
                                            TileSlot* slots, unsigned long long* ticket, cudaStream_t st) {
    (void)ticket;
    return exclusive_scan_launch_t<uint32_t>(in, n, cols, in_stride, out, out_stride, slots, st);
}
cudaError_t exclusive_scan_launch_u8(const uint8_t* in, int64_t n, int64_t* out, TileSlot* slots, unsigned long long* ticket,
                                     cudaStream_t st) {
    (void)ticket;
    return exclusive_scan_launch_t<uint8_t>(in, n, 1, 0, out, 0, slots, st);
}

__global__ void select_rows_kernel(const uint8_t* __restrict__ pass, const int64_t* __restrict__ off, int64_t n,
                                   int64_t* __restrict__ sel) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        if (pass[r]) sel[off[r]] = r;
}
cudaError_t select_rows_launch(const uint8_t* pass, const int64_t* off, int64_t n, int64_t* sel, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    select_rows_kernel<<<blocks, 256, 0, st>>>(pass, off, n, sel);
    return cudaGetLastError();
}

// ================================================================= FASTQ field extents
template <typename OffT>
struct FqLines {
    const uint8_t* buf;
    const OffT* line_end;
    int64_t begin, n;
    bool probe_cr;  // false: the scan saw no CR LF line ending in this range, so no line end needs the byte before it
    // start of line g and its end with a trailing CR removed
    __device__ __forceinline__ int64_t start(int64_t g) const { return g == 0 ? begin : (int64_t)line_end[g - 1] + 1; }
    __device__ __forceinline__ int64_t end(int64_t g, int64_t s) const {
        int64_t e = (int64_t)line_end[g];
        // e == n is the virtual terminator of an unterminated last line: it strips no CR
        if (probe_cr && e > s && e < n && buf[e - 1] == '\r') e--;
        return e;
    }
};

// Position of the first ' ' in [hs, e0), or e0.  A thread reads its header in 64-byte windows of four ALIGNED 16-byte
// loads issued together (a byte-at-a-time walk is ~35 dependent loads for an Illumina header), builds the 64-bit
// equality mask with the tile kernels' SWAR + IDP.4A helpers and takes the first set bit inside the range.  The last
// window may read up to 15 bytes past e0 <= n: inside the 64-byte slack every input buffer carries.
__device__ __forceinline__ int64_t first_space(const uint8_t* __restrict__ buf, int64_t hs, int64_t e0) {
    const uint32_t c7f = 0x7F7F7F7Fu, pat = 0x20202020u;
    int64_t base = hs - (int64_t)((uintptr_t)(buf + hs) & 15);
    while (base < e0) {
        uint64_t m = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (base + 16 * k < e0) {
                const uint4 v = *reinterpret_cast<const uint4*>(buf + base + 16 * k);
                m |= (uint64_t)nl_mask16r(v, c7f, pat) << (16 * k);
            }
        }
        if (hs > base) m &= ~0ull << (int)(hs - base);
        if (e0 - base < 64) m &= (1ull << (int)(e0 - base)) - 1ull;
        if (m) return base + (__ffsll((long long)m) - 1);
        base += 64;
    }
    return e0;
}

template <typename OffT>
__global__ void __launch_bounds__(256) fastq_fields_kernel(FqLines<OffT> L, const int64_t* __restrict__ sel, int64_t n_rows,
                                                           uint32_t* __restrict__ lens, uint8_t* __restrict__ desc_valid,
                                                           int64_t* __restrict__ starts, const ScanResult* __restrict__ scan) {
    // three scattered one-byte probes per record (a 32-byte sector each) are ~40 % of this kernel's traffic; K2 of the
    // scan counted the CR LF line endings of the range, and almost every file has none
    L.probe_cr = !scan || scan->crlf_lines != 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = sel ? sel[i] : i;
        const int64_t g = 4 * r;
        int64_t s0 = L.start(g), e0 = L.end(g, s0);
        int64_t hs = s0 + 1;  // skip '@'
        const int64_t sp = first_space(L.buf, hs, e0);  // first SPACE splits name / description
        uint32_t name_len = (uint32_t)(sp - hs);
        uint32_t desc_len = sp < e0 ? (uint32_t)(e0 - sp - 1) : 0u;
        int64_t s1 = (int64_t)L.line_end[g] + 1, e1 = L.end(g + 1, s1);
        int64_t s3 = (int64_t)L.line_end[g + 2] + 1, e3 = L.end(g + 3, s3);
        lens[0 * n_rows + i] = name_len;
        lens[1 * n_rows + i] = desc_len;
        lens[2 * n_rows + i] = (uint32_t)(e1 - s1);
        lens[3 * n_rows + i] = (uint32_t)(e3 - s3);
        desc_valid[i] = desc_len > 0;  // exon's builder maps an empty description to NULL
        if (starts) {
            starts[0 * n_rows + i] = hs;
            starts[1 * n_rows + i] = sp < e0 ? sp + 1 : e0;
            starts[2 * n_rows + i] = s1;
            starts[3 * n_rows + i] = s3;
        }
    }
}

// ================================================================= gathers
// One warp per row: copies len bytes from src to dst.  dst-aligned 4-byte words
// in the body (two aligned source loads + funnel shift), bytes at the edges.
__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int64_t len, int lane) {
    if (len <= 0) return;
    int head = (int)((4 - ((uintptr_t)dst & 3)) & 3);
    if (head > len) head = (int)len;
    if (lane < head) dst[lane] = src[lane];
    dst += head;
    src += head;
    len -= head;
    const int64_t nwords = len >> 2;
    const int sh = (int)((uintptr_t)src & 3);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(src - sh);
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    if (sh == 0) {
        for (int64_t w = lane; w < nwords; w += 32) dw[w] = sw[w];
    } else {
        for (int64_t w = lane; w < nwords; w += 32) dw[w] = __funnelshift_r(sw[w], sw[w + 1], sh * 8);
    }
    const int tail = (int)(len & 3);
    if (lane < tail) dst[nwords * 4 + lane] = src[nwords * 4 + lane];
}

// Output-centric gather: the output column is cut into spans of GS_SPAN bytes, one per block; a thread owns 16-byte
// (dst-aligned) chunks of the span.  The rows that cover the span are found once per block (binary search of the
// offsets) and their offsets / source addresses are held in shared memory GS_ROWS at a time, so a chunk costs a
// ~9-step shared-memory search, two aligned 16-byte loads funnel-shifted to the destination alignment and ONE
// 16-byte store; chunks that straddle rows fall back to bytes.  Throughput does not depend on the row length: 12-byte
// descriptions, 150-byte reads and 250 Mbp contigs all stream.  `off` = exclusive prefix sum of the row lengths,
// n_rows + 1 entries (monotone; empty rows are fine).
constexpr int GS_SPAN = 65536, GS_ROWS = 512, GS_THREADS = 256, GS_SEG = 64;

struct SrcRanges {  // generic: start[i]
    const int64_t* start;
    __device__ __forceinline__ int64_t operator()(int64_t i) const { return start[i]; }
};
template <typename OffT>
struct SrcFastq {  // field `col` of the selected FASTQ records
    FqLines<OffT> L;
    const int64_t* sel;
    const uint32_t* name_len;
    int col;
    __device__ __forceinline__ int64_t operator()(int64_t i) const {
        const int64_t g = 4 * (sel ? sel[i] : i);
        if (col == 0) return L.start(g) + 1;
        if (col == 1) return L.start(g) + 1 + name_len[i] + 1;
        if (col == 2) return (int64_t)L.line_end[g] + 1;
        return (int64_t)L.line_end[g + 2] + 1;
    }
};

// 16 bytes from an arbitrary address: five 4-byte aligned words (they share at most two 32-byte sectors, so the extra
// load instructions hit L1) and four funnel shifts.  An earlier version used two aligned 16-byte loads and a select
// network to pick the five words: 15 SEL + 6 ISETP per chunk, the hottest lines of the kernel.
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t* __restrict__ src) {
    const int bs = (int)((uintptr_t)src & 3) * 8;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src - (bs >> 3));
    // no branch on bs: the lanes of a warp mix aligned and unaligned sources, and funnelshift_r(x, y, 0) = x (the fifth
    // word is then read and ignored: 4 bytes into the slack every input buffer carries)
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}

// 128-bit byte mask as four words: bytes [0, s) of a 16-byte chunk, 0 < s < 16 (32-bit arithmetic only)
__device__ __forceinline__ uint32_t bytes_below32(int k) { return k >= 4 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << (8 * k)) - 1u)); }

// Positions inside a span are kept RELATIVE to span_lo (the 16-byte aligned output address the span starts at), so the
// inner loop runs on 32-bit integers: s_rel[i] = off[row0 + i] - span_lo clamped to +-2^30 (rows that begin before the
// span compare as "before", which is all the search needs), s_adj[i] = offset in buf of the byte that lands on relative
// position 0 if row i is extended backwards (64-bit, one add per piece).
// A 16-byte chunk is assembled from the pieces of the rows that meet in it -- one unaligned 16-byte load + a byte mask
// per piece -- and stored ONCE.  (The first version copied straddling chunks byte by byte: one chunk in ten for
// 150-byte rows, but under warp divergence 2/3 of all issued instructions.)
__device__ __forceinline__ void fill_map_lut_fwd(uint8_t* lut, int mode);
// kMap: the gathered bytes go through the reverse_complement / complement LUT (sequence_functions/module.cpp:30-121) on
// their way out -- `SELECT reverse_complement(sequence) FROM read_fastq(...)` on a device-resident file is then ONE pass
// over the sequence bytes instead of gather + map.  *bad = min over invalid bytes of (output position << 8 | byte).
template <typename SrcFn, bool kMap>
__global__ void __launch_bounds__(GS_THREADS, 8) gather_span_kernel(const uint8_t* __restrict__ buf, SrcFn srcfn, const int64_t* __restrict__ off,
                                                                 int64_t n_rows, uint8_t* __restrict__ out, int mode, unsigned long long* bad_out) {
    __shared__ int32_t s_rel[GS_ROWS + 1];
    __shared__ int64_t s_adj[GS_ROWS];
    __shared__ uint16_t s_seg[GS_SPAN / GS_SEG];  // row (batch-local) that holds relative position GS_SEG * j
    __shared__ uint8_t s_lut[kMap ? 256 : 1];
    const int t = threadIdx.x;
    unsigned long long bad = ~0ull;
    if (kMap) {
        s_lut[t] = 0;  // GS_THREADS == 256 entries
        __syncthreads();
        if (t == 0) fill_map_lut_fwd(s_lut, mode);
        __syncthreads();
    }
    auto map1 = [&](uint32_t c, int64_t pos) -> uint32_t {
        const uint32_t m = s_lut[c];
        if (m == 0) {
            const unsigned long long w = ((unsigned long long)pos << 8) | c;
            bad = w < bad ? w : bad;
        }
        return m;
    };
    auto map4 = [&](uint32_t x, int64_t pos) -> uint32_t {  // one validity test per word (see fastq_split_kernel)
        const uint32_t m = (uint32_t)s_lut[x & 0xFF] | ((uint32_t)s_lut[(x >> 8) & 0xFF] << 8) | ((uint32_t)s_lut[(x >> 16) & 0xFF] << 16) |
                           ((uint32_t)s_lut[x >> 24] << 24);
        if (((m - 0x01010101u) & ~m & 0x80808080u) != 0u) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (((m >> (8 * k)) & 0xFFu) == 0u) {
                    const unsigned long long w = ((unsigned long long)(pos + k) << 8) | ((x >> (8 * k)) & 0xFFu);
                    bad = w < bad ? w : bad;
                }
        }
        return m;
    };
    const int64_t total = off[n_rows];
    // spans are aligned to 16 bytes of the OUTPUT ADDRESS so that chunk stores are aligned whatever `out` is
    const int mis = (int)((uintptr_t)out & 15);
    // persistent blocks: the column's size is known only on the device, so the grid is sized for the machine and
    // every block walks the spans blockIdx.x, blockIdx.x + gridDim.x, ... (all loop bounds are block-uniform)
    for (int64_t span = blockIdx.x;; span += gridDim.x) {
        const int64_t span_lo = span * GS_SPAN - mis;  // negative for span 0 when out is misaligned
        if (span_lo >= total) break;
        const int64_t lo = span_lo < 0 ? 0 : span_lo;
        const int64_t hi = span_lo + GS_SPAN < total ? span_lo + GS_SPAN : total;
        if (lo >= hi) continue;
        // last row r with off[r] <= lo: a GS_THREADS-ary search by the whole block (3 rounds of one global load each
        // for 16 M rows)
        int64_t row0;
        {
            int64_t a = 0, b = n_rows;  // invariant: off[a] <= lo < off[b] (off[n_rows] = total > lo)
            while (b - a > 1) {
                const int64_t step = (b - a + GS_THREADS - 1) / GS_THREADS;
                const int64_t m = a + (int64_t)(t + 1) * step;
                const int c = __syncthreads_count(m < b && off[m] <= lo);  // off is monotone: the true probes are a prefix
                const int64_t na = a + (int64_t)c * step;
                b = na + step < b ? na + step : b;
                a = na;
            }
            row0 = a;
        }
        int lo_rel = (int)(lo - span_lo);
        const int hi_rel = (int)(hi - span_lo);
        uint8_t* const span_out = out + span_lo;  // 16-byte aligned
        while (lo_rel < hi_rel) {
            const int64_t left = n_rows - row0;
            const int cap = left < GS_ROWS ? (int)left : GS_ROWS;
            __syncthreads();
            for (int i = t; i <= cap; i += GS_THREADS) {
                int64_t r = off[row0 + i] - span_lo;
                r = r < -(1 << 30) ? -(1 << 30) : (r > (1 << 30) ? (1 << 30) : r);
                s_rel[i] = (int32_t)r;
            }
            __syncthreads();
            // rows that begin before `hi` are the ones this span needs (s_rel[0] <= lo_rel < hi_rel, so at least one);
            // only those pay for their source address (SrcFastq: two dependent scattered loads per row)
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < GS_ROWS / GS_THREADS; k++) {
                const int i = t + k * GS_THREADS;
                cnt += __syncthreads_count(i < cap && s_rel[i] < hi_rel);
            }
            for (int i = t; i < cnt; i += GS_THREADS) s_adj[i] = srcfn(row0 + i) - (off[row0 + i] - span_lo);
            // one binary search per 64-byte segment (one per thread) instead of one per 16-byte chunk
            const int seg_end = ((s_rel[cnt] < hi_rel ? s_rel[cnt] : hi_rel) + GS_SEG - 1) / GS_SEG;  // segments this batch serves
            for (int j = lo_rel / GS_SEG + t; j < seg_end; j += GS_THREADS) {
                const int q = j * GS_SEG;
                int a = 0, b = cnt;  // last i in [0, cnt) with s_rel[i] <= q (s_rel[0] <= lo_rel; positions before lo_rel are never looked up)
                while (b - a > 1) {
                    const int m = (a + b) >> 1;
                    if (s_rel[m] <= q) a = m;
                    else b = m;
                }
                s_seg[j] = (uint16_t)a;
            }
            __syncthreads();
            // rows row0 .. row0 + cnt cover [lo_rel, batch_hi)
            const int batch_hi = s_rel[cnt] < hi_rel ? s_rel[cnt] : hi_rel;
            // One chunk, general case: pieces of several rows (or a ragged edge of the column / batch).
            auto chunk_general = [&](int c0, int a) {
                const int q0 = c0 < lo_rel ? lo_rel : c0, q1 = c0 + 16 > batch_hi ? batch_hi : c0 + 16;
                bool whole = q1 - q0 == 16;
                uint4 acc = make_uint4(0, 0, 0, 0);
                if (whole) {
                    int q = q0;
                    while (q < q1) {
                        while (s_rel[a + 1] <= q) a++;  // skips empty rows; q < batch_hi <= s_rel[cnt] bounds it
                        const int e = s_rel[a + 1] < q1 ? s_rel[a + 1] : q1;
                        const int64_t src = s_adj[a] + c0;  // source of the chunk's byte 0 if row a reached that far back
                        if (src < 0) {  // would read in front of the buffer (first bytes of the input): take the byte path
                            whole = false;
                            break;
                        }
                        const uint4 v = load16_unaligned(buf + src);
                        if (q == c0) {
                            acc = v;  // bytes past this row's end are garbage until the next piece overwrites them
                        } else {      // pieces arrive in order and the last one ends the chunk: keep [0, s), take [s, 16)
                            const int sb = q - c0;
                            const uint32_t m0 = bytes_below32(sb), m1 = bytes_below32(sb - 4), m2 = bytes_below32(sb - 8), m3 = bytes_below32(sb - 12);
                            acc.x = (acc.x & m0) | (v.x & ~m0);
                            acc.y = (acc.y & m1) | (v.y & ~m1);
                            acc.z = (acc.z & m2) | (v.z & ~m2);
                            acc.w = (acc.w & m3) | (v.w & ~m3);
                        }
                        q = e;
                    }
                }
                if (whole) {
                    if (kMap) {
                        const int64_t p = span_lo + c0;
                        acc.x = map4(acc.x, p), acc.y = map4(acc.y, p + 4), acc.z = map4(acc.z, p + 8), acc.w = map4(acc.w, p + 12);
                    }
                    *reinterpret_cast<uint4*>(span_out + c0) = acc;
                } else {  // ragged first / last chunk of the column or of a row batch
                    for (int q = q0; q < q1; q++) {
                        while (s_rel[a + 1] <= q) a++;
                        const uint8_t b = buf[s_adj[a] + q];
                        span_out[q] = kMap ? (uint8_t)map1(b, span_lo + q) : b;
                    }
                }
            };
            // local row of a position: last i with s_rel[i] <= q -- from the segment's row, a step or two forward
            auto row_of = [&](int q) {
                int a = s_seg[q / GS_SEG];
                while (a + 1 < cnt && s_rel[a + 1] <= q) a++;
                return a;
            };
            // One chunk per thread and iteration at 8 CTAs / SM.  (Two chunks in flight per thread at 6 CTAs / SM -- the
            // register cost of the second set of loads -- measured slower: 556 vs 452 us per 600 MB column.)
            for (int k = (lo_rel >> 4) + t; (k << 4) < batch_hi; k += GS_THREADS) {
                const int c0 = k << 4;
                const int q0 = c0 < lo_rel ? lo_rel : c0;
                // every chunk takes the piece loop; a separate "inside one row: load + store" branch in front of it
                // measured 14 % slower on the same box (455 vs 518 us; scripts/gpu_ab.sh)
                chunk_general(c0, row_of(q0));
            }
            lo_rel = batch_hi;
            row0 += cnt;  // the batch ended on its last row (lo = off[row0 + cnt]), or lo == hi and the loop ends
        }
    }
    if (kMap && bad != ~0ull) atomicMin(bad_out, bad);
}

template <typename SrcFn>
static cudaError_t gather_span_launch(const uint8_t* buf, SrcFn fn, const int64_t* off, int64_t n_rows, int64_t total_hint, uint8_t* out,
                                      cudaStream_t st, int mode = 0, unsigned long long* bad = nullptr) {
    if (bad) {  // map requested (mode 0 is a valid map): the invalid-byte word starts at "none"
        cudaError_t e = cudaMemsetAsync(bad, 0xFF, sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
    }
    if (n_rows == 0) return cudaSuccess;
    // off[n_rows] is known only on the device; total_hint (an upper bound) merely keeps tiny columns from launching a full grid
    int64_t blocks = (total_hint + 15 + GS_SPAN - 1) / GS_SPAN + 1;
    const int64_t machine = 148 * 8;  // 8 resident blocks of 256 threads per SM
    if (blocks > machine) blocks = machine;
    if (blocks < 1) blocks = 1;
    if (bad) gather_span_kernel<SrcFn, true><<<(unsigned)blocks, GS_THREADS, 0, st>>>(buf, fn, off, n_rows, out, mode, bad);
    else gather_span_kernel<SrcFn, false><<<(unsigned)blocks, GS_THREADS, 0, st>>>(buf, fn, off, n_rows, out, 0, nullptr);
    return cudaGetLastError();
}

// ================================================================= fused column split (FASTQ, all rows)
// ONE launch for what used to be the field split, the offset scan and four column gathers: a block takes 256
// consecutive records (ticket order), works out their field extents, scans the four field lengths inside the block,
// publishes the block totals, obtains its output offsets with a decoupled look-back over the earlier blocks, writes the
// Arrow offsets of its rows and then copies all four columns of its rows.  The four gathers of a block read one
// contiguous ~90 KB window of the input while it is hot in L1 / L2, so every input DRAM atom is fetched once per pass
// (the per-column gathers fetched the file about three times: profiles/r02r_kernel_traffic.txt), and the header bytes
// the field split reads are the ones the name / description copies need a moment later.
// Output layout = a6 / a8 of SURVEY 8: per column int64 offsets[n_rows + 1] (exclusive prefix of the lengths) + bytes.
constexpr int SP_ROWS = 256, SP_THREADS = 256;
struct alignas(128) SplitDesc {  // one per block of rows: its totals, then its inclusive prefix
    unsigned long long agg[4];
    unsigned long long incl[4];
    uint32_t status;  // 0 = nothing yet, 1 = agg valid, 2 = incl valid
    uint32_t pad[15];
};
static_assert(sizeof(SplitDesc) == 128, "SplitDesc layout");

struct SplitArgs {
    int64_t n_rows;
    uint32_t mask;            // bit c: column c is copied (its offsets are always produced)
    int64_t* off;             // [4][n_rows + 1]
    uint8_t* desc_valid;      // [n_rows]
    uint8_t* out[4];          // 16-byte aligned
    int64_t cap[4];
    SplitDesc* desc;          // [ceil(n_rows / SP_ROWS)] zeroed
    unsigned int* ticket;     // zeroed
    unsigned int* overflow;   // set to 1 if a column does not fit its capacity
    const ScanResult* scan;   // optional: CR LF knowledge of the scan (see fastq_fields_kernel)
    int map_mode;             // kMap: EXB_MAP_* applied to column 2 (sequence)
    int rows_per_block;       // <= SP_ROWS: long records get fewer rows per block so that the grid stays full
    int win_bytes;            // bytes of input a block may stage in (dynamic) shared memory; 0 = read global memory
    int prefetch;             // global path: prefetch the block's input window into L2 before the field split
    unsigned long long* bad;  // kMap: min over invalid bytes of (output position << 8 | byte)
};

template <typename OffT, bool kMap>
__global__ void __launch_bounds__(SP_THREADS, 5) fastq_split_kernel(FqLines<OffT> L, const SplitArgs a) {
    __shared__ int64_t s_start[4][SP_ROWS];       // source offset of each field
    __shared__ int64_t s_loc[4][SP_ROWS + 1];     // block-local exclusive offsets, closed with the block totals
    __shared__ unsigned long long s_warp[4][SP_THREADS / 32];
    __shared__ unsigned long long s_wmax[4][SP_THREADS / 32];
    __shared__ unsigned long long s_base[4];
    __shared__ unsigned long long s_maxlen[4];   // longest field of the block, per column
    __shared__ uint4 s_below[17];                // s_below[k]: 0xFF in the first k bytes of a 16-byte chunk
    __shared__ unsigned int s_blk;
    __shared__ int64_t s_w[2];                   // the block's input window [w0, w1)
    extern __shared__ __align__(16) uint8_t s_win[];  // a.win_bytes + 64
    __shared__ uint8_t s_lut[kMap ? 256 : 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned long long bad = ~0ull;
    if (kMap) {
        s_lut[t] = 0;
        __syncthreads();
        if (t == 0) fill_map_lut_fwd(s_lut, a.map_mode);
        __syncthreads();
    }
    if (t < 17) s_below[t] = make_uint4(bytes_below32(t), bytes_below32(t - 4), bytes_below32(t - 8), bytes_below32(t - 12));
    L.probe_cr = !a.scan || a.scan->crlf_lines != 0;
    const bool want_header = (a.mask & 3u) != 0;
    const uint8_t* __restrict__ buf = L.buf;
    const int RPB = a.rows_per_block;
    const int64_t n_blocks = (a.n_rows + RPB - 1) / RPB;
    for (;;) {
        __syncthreads();  // the previous iteration's readers of the shared arrays are done
        if (t == 0) s_blk = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const int64_t blk = s_blk;
        if (blk >= n_blocks) break;
        const int64_t r = blk * RPB + t;
        const int rows_here = (int)((a.n_rows - blk * RPB) < RPB ? (a.n_rows - blk * RPB) : RPB);
        const bool active = t < rows_here;
        // ---- the block's input window: the bytes of its records are contiguous in the file.  When they fit, they are staged
        // in shared memory with one round of coalesced 16-byte asynchronous copies (every thread has several in flight) and
        // everything below -- header search, CR probes, the copies of all four columns -- reads them from there: the
        // unaligned source reads of the gathers then cost a shared-memory access instead of a global-memory latency
        // (a version that read the input straight from global memory was latency bound at 48 % issue utilisation and
        // 34 % of DRAM bandwidth: profiles/round2_split_v2_ncu.txt).  Long records (ONT) keep the global path.
        if (t == 0) s_w[0] = L.start(4 * (blk * RPB));
        if (t == 32) s_w[1] = (int64_t)L.line_end[4 * (blk * RPB + rows_here) - 1] + 1;
        __syncthreads();
        const int64_t w0a = s_w[0] & ~(int64_t)15, w1 = s_w[1] > L.n ? L.n : s_w[1];
        const bool use_win = w1 - w0a <= a.win_bytes;
        if (use_win) {
            const int chunks = (int)((w1 - w0a + 15 + 32) >> 4);  // 32 bytes of look-ahead for the 16-byte reads at a row's end
            for (int k = t; k < chunks; k += SP_THREADS) cp_async16(s_win + 16 * k, buf + w0a + 16 * (int64_t)k, 16);
            cp_async_commit();
        }
        else if (a.prefetch && (a.mask & 0xFu) == 0xFu) {  // the block's window will be read piecemeal below: start its trip from
            // DRAM to L2 now -- when every column is copied; with a projection the window holds bytes nobody reads (the quality
            // lines of a `sequence`-only split were fetched for nothing: 3.2 GB read for a 1.2 GB input, half of it useful)
            for (int64_t o = w0a + 128 * (int64_t)t; o < w1; o += 128 * SP_THREADS) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(buf + o));
        }
        // (the input carries 64 bytes of slack behind n, so the look-ahead never leaves the allocation)
        const uint8_t* __restrict__ rd = use_win ? (const uint8_t*)s_win - w0a : buf;  // rd[file offset] = that byte
        const int64_t lo_guard = use_win ? w0a + 4 : 4;  // an unaligned 16-byte read must not start below this offset
        // ---- A. field extents of this thread's record
        unsigned long long len[4] = {0, 0, 0, 0};
        bool valid_desc = false;
        int64_t le[5] = {0, 0, 0, 0, 0};
        if (active) {
            const int64_t g = 4 * r;
            le[0] = g == 0 ? L.begin - 1 : (int64_t)L.line_end[g - 1];
#pragma unroll
            for (int k = 0; k < 4; k++) le[k + 1] = (int64_t)L.line_end[g + k];
        }
        if (use_win) {
            cp_async_wait<0>();
            __syncthreads();
        }
        if (active) {
            // start of line k = le[k] + 1; its end with a trailing CR removed (e == n is the virtual terminator: no CR)
            auto line_end_of = [&](int k) {
                int64_t e = le[k + 1];
                const int64_t sl = le[k] + 1;
                if (L.probe_cr && e > sl && e < L.n && rd[e - 1] == '\r') e--;
                return e;
            };
            const int64_t s0 = le[0] + 1, e0 = line_end_of(0);
            const int64_t hs = s0 + 1;  // skip '@'
            int64_t sp = e0;
            if (want_header) sp = first_space(rd, hs, e0);  // first SPACE splits name / description
            const int64_t s1 = le[1] + 1, e1 = line_end_of(1);
            const int64_t s3 = le[3] + 1, e3 = line_end_of(3);
            len[0] = (unsigned long long)(sp - hs);
            len[1] = sp < e0 ? (unsigned long long)(e0 - sp - 1) : 0ull;
            len[2] = (unsigned long long)(e1 - s1);
            len[3] = (unsigned long long)(e3 - s3);
            valid_desc = len[1] > 0;  // exon's builder maps an empty description to NULL
            s_start[0][t] = hs;
            s_start[1][t] = sp < e0 ? sp + 1 : e0;
            s_start[2][t] = s1;
            s_start[3][t] = s3;
        }
        // ---- block-wide exclusive scan of the four lengths
        unsigned long long incl[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            incl[c] = warp_incl_scan_u64(len[c]);
            if (lane == 31) s_warp[c][warp] = incl[c];
            unsigned long long mx = len[c];
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, dd);
                mx = o > mx ? o : mx;
            }
            if (lane == 0) s_wmax[c][warp] = mx;
        }
        __syncthreads();
        unsigned long long tot[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            unsigned long long before = 0, all = 0;
#pragma unroll
            for (int w = 0; w < SP_THREADS / 32; w++) {
                const unsigned long long v = s_warp[c][w];
                before += w < warp ? v : 0ull;
                all += v;
            }
            tot[c] = all;
            s_loc[c][t] = (int64_t)(before + incl[c] - len[c]);  // threads past the block's rows write the total
            if (t == 0) {
                s_loc[c][SP_ROWS] = (int64_t)all;
                unsigned long long mx = 0;
#pragma unroll
                for (int w = 0; w < SP_THREADS / 32; w++) mx = s_wmax[c][w] > mx ? s_wmax[c][w] : mx;
                s_maxlen[c] = mx;
            }
        }
        // ---- publish the totals, look back for the block's base (warp 0)
        if (warp == 0) {
            SplitDesc* d = a.desc + blk;
            if (lane < 4) d->agg[lane] = tot[lane];
            if (blk == 0 && lane < 4) d->incl[lane] = tot[lane];
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u32(&d->status, blk == 0 ? 2u : 1u);
            unsigned long long ex[4] = {0, 0, 0, 0};
            if (blk > 0) {
                int64_t idx = blk - 1;
                for (;;) {
                    const int64_t p = idx - lane;
                    uint32_t st = 2u;
                    if (p >= 0) {
                        st = ld_acquire_u32(&a.desc[p].status);
                        while (st == 0u) st = ld_acquire_u32(&a.desc[p].status);  // an earlier ticket: it is running and will publish
                    }
                    const uint32_t has_incl = __ballot_sync(0xffffffffu, st == 2u);
                    const int first = has_incl ? __ffs((int)has_incl) - 1 : 32;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        unsigned long long v = 0;
                        if (p >= 0 && lane <= first) v = lane < first ? __ldcg(&a.desc[p].agg[c]) : __ldcg(&a.desc[p].incl[c]);  // written by another SM: L2
#pragma unroll
                        for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_xor_sync(0xffffffffu, v, dd);
                        ex[c] += v;
                    }
                    if (has_incl) break;
                    idx -= 32;
                }
                if (lane < 4) d->incl[lane] = ex[lane] + tot[lane];
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_u32(&d->status, 2u);
            }
            if (lane < 4) s_base[lane] = ex[lane];
        }
        __syncthreads();
        // ---- B. Arrow offsets and validity of the block's rows
        if (active) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int64_t* o = a.off + (int64_t)c * (a.n_rows + 1);
                o[r] = (int64_t)s_base[c] + s_loc[c][t];
                if (r == a.n_rows - 1) o[a.n_rows] = (int64_t)s_base[c] + s_loc[c][t] + (int64_t)len[c];
            }
            a.desc_valid[r] = valid_desc ? 1 : 0;
        }
        // ---- C. the four columns of the block's rows.
        // Columns of long rows (sequence, quality_scores; ONT reads) are copied ROW-centric: a group of lanes owns a row,
        // each lane moves whole destination-aligned 16-byte chunks of it (one unaligned 16-byte load, one aligned store) and
        // then the <= 30 ragged bytes at the row's two ends one by one -- ~20 warp instructions per 150-byte row.
        // Columns of short rows (name, description) are copied CHUNK-centric: 16-byte destination-aligned chunks assembled
        // from the pieces of the rows that meet in them (see gather_span_kernel); the ragged bytes at the two ends of the
        // block's range are written one by one (the neighbouring blocks write the rest of those chunks).
        // (A chunk-centric copy of all four columns executed 250 thread instructions per chunk, 1.0 G warp instructions for
        // a 1.4 GB file, and was issue bound at 1.5 ms: profiles/round2_split_v1_ncu.txt.)
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            if (!((a.mask >> c) & 1u)) continue;
            const int64_t B0 = (int64_t)s_base[c], B1 = B0 + s_loc[c][SP_ROWS];
            if (B1 > a.cap[c]) {
                if (t == 0) *a.overflow = 1u;
                continue;
            }
            if (B1 == B0) continue;
            uint8_t* __restrict__ out = a.out[c];
            const int64_t* __restrict__ loc = s_loc[c];
            const int64_t* __restrict__ src0 = s_start[c];
            const bool map = kMap && c == 2;
            auto map1 = [&](uint32_t ch, int64_t pos) -> uint32_t {
                const uint32_t m = s_lut[ch];
                if (m == 0) {
                    const unsigned long long w = ((unsigned long long)pos << 8) | ch;
                    bad = w < bad ? w : bad;
                }
                return m;
            };
            auto map4 = [&](uint32_t x, int64_t pos) -> uint32_t {
                // four table look-ups, ONE validity test per word: the table holds ASCII letters or 0 (= invalid), so the
                // SWAR zero-byte test is exact; the per-byte search only runs for a word that holds an invalid character
                // (a test and a branch per byte were 3/4 of this kernel's instructions on ONT reads: profiles/round2_split_map_ncu.txt)
                const uint32_t m = (uint32_t)s_lut[x & 0xFF] | ((uint32_t)s_lut[(x >> 8) & 0xFF] << 8) | ((uint32_t)s_lut[(x >> 16) & 0xFF] << 16) |
                                   ((uint32_t)s_lut[x >> 24] << 24);
                if (((m - 0x01010101u) & ~m & 0x80808080u) != 0u) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (((m >> (8 * k)) & 0xFFu) == 0u) {
                            const unsigned long long w = ((unsigned long long)(pos + k) << 8) | ((x >> (8 * k)) & 0xFFu);
                            bad = w < bad ? w : bad;
                        }
                }
                return m;
            };
            const int64_t col_bytes = B1 - B0;
            const int max_len = (int)(s_maxlen[c] > 0x7FFFFFFFull ? 0x7FFFFFFFull : s_maxlen[c]);
            if (col_bytes >= 48 * (int64_t)rows_here) {
                // ---- row-centric
                const int per_row = (max_len >> 4) + 1;                       // lanes a row can keep busy with whole chunks
                const int rpi = per_row >= 32 ? 1 : (32 / per_row > 8 ? 8 : 32 / per_row);  // rows per warp iteration
                const int slot = 32 / rpi;
                const int g = lane / slot, k = lane - g * slot;
                for (int base = warp * rpi; base < rows_here; base += (SP_THREADS / 32) * rpi) {
                    const int i = base + g;
                    if (g >= rpi || i >= rows_here) continue;
                    const int64_t D0 = B0 + loc[i], D1 = B0 + loc[i + 1];
                    const int64_t S0 = src0[i] - D0;  // source = S0 + destination
                    int64_t a0 = (D0 + 15) & ~(int64_t)15, a1 = D1 & ~(int64_t)15;
                    if (a0 > a1) a0 = a1 = D1;        // the row holds no whole chunk: all of it is "head"
                    if (S0 + a0 < lo_guard) a0 = a1 = D1;  // an unaligned load would start in front of the buffer / window: bytes
                    const int64_t n_ch = (a1 - a0) >> 4;
                    for (int64_t kk = k; kk < n_ch; kk += slot) {
                        const int64_t d = a0 + 16 * kk;
                        uint4 v = load16_unaligned(rd + S0 + d);
                        if (map) v.x = map4(v.x, d), v.y = map4(v.y, d + 4), v.z = map4(v.z, d + 8), v.w = map4(v.w, d + 12);
                        *reinterpret_cast<uint4*>(out + d) = v;
                    }
                    const int head_n = (int)(a0 - D0), rag = head_n + (int)(D1 - a1);
                    for (int bb = k; bb < rag; bb += slot) {
                        const int64_t d = bb < head_n ? D0 + bb : a1 + (bb - head_n);
                        const uint8_t ch = rd[S0 + d];
                        out[d] = map ? (uint8_t)map1(ch, d) : ch;
                    }
                }
                continue;
            }
            // ---- chunk-centric
            // block-local row of a block-local position q: last i < rows_here with loc[i] <= q.  Rows of a column are about
            // the same length, so the proportional guess is a step or two away.
            const float inv = (float)rows_here / (float)col_bytes;
            auto row_of = [&](int64_t q) {
                int i = (int)((float)q * inv);
                i = i < rows_here - 1 ? i : rows_here - 1;
                while (loc[i] > q) i--;
                while (i + 1 < rows_here && loc[i + 1] <= q) i++;
                return i;
            };
            const int64_t A0 = (B0 + 15) & ~(int64_t)15, A1 = B1 & ~(int64_t)15;  // whole chunks: [A0, A1)
            auto bytes = [&](int64_t from, int64_t to) {  // ragged ends, at most 15 bytes each
                for (int64_t q = from + t; q < to; q += SP_THREADS) {
                    const int i = row_of(q - B0);
                    const uint8_t b = rd[src0[i] + (q - B0 - loc[i])];
                    out[q] = map ? (uint8_t)map1(b, q) : b;
                }
            };
            if (A0 >= A1) {  // the block's range does not hold a whole chunk
                bytes(B0, B1);
                continue;
            }
            bytes(B0, A0);
            bytes(A1, B1);
            for (int64_t c0 = A0 + 16 * (int64_t)t; c0 < A1; c0 += 16 * SP_THREADS) {
                int i = row_of(c0 - B0);
                int64_t q = c0;
                const int64_t q1 = c0 + 16;
                uint4 acc = make_uint4(0, 0, 0, 0);
                bool whole = true;
                while (q < q1) {
                    while (loc[i + 1] <= q - B0) i++;  // skips empty rows; q < B1 bounds it
                    const int64_t row_end = B0 + loc[i + 1];
                    const int64_t e = row_end < q1 ? row_end : q1;
                    const int64_t src = src0[i] + (c0 - B0 - loc[i]);  // source of the chunk's byte 0 if row i reached that far back
                    if (src < lo_guard) {  // would read in front of the buffer / window: bytes
                        whole = false;
                        break;
                    }
                    const uint4 v = load16_unaligned(rd + src);
                    if (q == c0) {
                        acc = v;  // bytes past this row's end are garbage until the next piece overwrites them
                    } else {      // pieces arrive in order: keep [0, s), take [s, 16)
                        const uint4 m = s_below[(int)(q - c0)];
                        acc.x = (acc.x & m.x) | (v.x & ~m.x);
                        acc.y = (acc.y & m.y) | (v.y & ~m.y);
                        acc.z = (acc.z & m.z) | (v.z & ~m.z);
                        acc.w = (acc.w & m.w) | (v.w & ~m.w);
                    }
                    q = e;
                }
                if (whole) {
                    if (map) acc.x = map4(acc.x, c0), acc.y = map4(acc.y, c0 + 4), acc.z = map4(acc.z, c0 + 8), acc.w = map4(acc.w, c0 + 12);
                    *reinterpret_cast<uint4*>(out + c0) = acc;
                } else {
                    for (int64_t qq = c0; qq < q1; qq++) {
                        const int k = row_of(qq - B0);
                        const uint8_t b = rd[src0[k] + (qq - B0 - loc[k])];
                        out[qq] = map ? (uint8_t)map1(b, qq) : b;
                    }
                }
            }
        }
    }
    if (kMap && bad != ~0ull) atomicMin(a.bad, bad);
}

constexpr int SP_MIN_ROWS = 8;
int64_t fastq_split_scratch_bytes(int64_t n_rows) {
    return 256 + ((n_rows + SP_MIN_ROWS - 1) / SP_MIN_ROWS + 1) * (int64_t)sizeof(SplitDesc);
}

template <typename OffT>
static cudaError_t split_launch_t(const uint8_t* buf, int64_t begin, int64_t n, const void* line_end, int64_t n_rows, uint32_t mask, int64_t* off,
                                  uint8_t* desc_valid, uint8_t* const* out, const int64_t* cap, void* scratch, const ScanResult* scan, int map_mode,
                                  unsigned long long* bad, cudaStream_t st) {
    SplitArgs a;
    a.n_rows = n_rows;
    a.mask = mask;
    a.off = off;
    a.desc_valid = desc_valid;
    for (int c = 0; c < 4; c++) {
        a.out[c] = out[c];
        a.cap[c] = cap[c];
    }
    uint8_t* s = reinterpret_cast<uint8_t*>(scratch);
    a.ticket = reinterpret_cast<unsigned int*>(s);
    a.overflow = reinterpret_cast<unsigned int*>(s + 8);
    a.desc = reinterpret_cast<SplitDesc*>(s + 256);
    a.scan = scan;
    a.map_mode = map_mode;
    a.bad = bad;
    // A block takes about 110 KB of input (256 Illumina records; long records get fewer rows so that the grid stays full).
    // EXB_SPLIT_WIN_KB > 0 stages each block's window in shared memory instead (fewer rows per block, sized to fit).
    static const int win_kb = getenv("EXB_SPLIT_WIN_KB") ? atoi(getenv("EXB_SPLIT_WIN_KB")) : 0;
    static const int prefetch = getenv("EXB_SPLIT_PREFETCH") ? atoi(getenv("EXB_SPLIT_PREFETCH")) : 1;
    const int64_t avg = n_rows > 0 ? (n - begin) / n_rows + 1 : 1;
    a.win_bytes = win_kb > 0 ? (win_kb << 10) : 0;
    a.prefetch = prefetch;
    int64_t rpb = a.win_bytes ? (int64_t)(a.win_bytes * 0.85) / avg : (112 << 10) / avg;
    rpb = rpb < SP_MIN_ROWS ? SP_MIN_ROWS : (rpb > SP_ROWS ? SP_ROWS : rpb);
    a.rows_per_block = (int)rpb;
    const int64_t n_blocks = (n_rows + rpb - 1) / rpb;
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)(256 + (n_blocks + 1) * (int64_t)sizeof(SplitDesc)), st);
    if (e != cudaSuccess) return e;
    if (bad && (e = cudaMemsetAsync(bad, 0xFF, 8, st)) != cudaSuccess) return e;
    FqLines<OffT> L{buf, reinterpret_cast<const OffT*>(line_end), begin, n, true};
    int64_t blocks = n_blocks;
    const int64_t machine = 148 * 5;  // 5 resident blocks of 256 threads per SM
    if (blocks > machine) blocks = machine;
    if (blocks < 1) blocks = 1;
    const size_t dyn = a.win_bytes ? (size_t)a.win_bytes + 64 : 16;
    if (dyn > 16) {  // beyond the 48 KB default (static arrays included) the kernel must opt in
        cudaFuncSetAttribute(fastq_split_kernel<OffT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        cudaFuncSetAttribute(fastq_split_kernel<OffT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    }
    int per_sm = 5;
    if (a.win_bytes) {
        per_sm = (int)((227 << 10) / ((int64_t)dyn + 18 * 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 5 ? 5 : per_sm);
    }
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    if (bad) fastq_split_kernel<OffT, true><<<(unsigned)blocks, SP_THREADS, dyn, st>>>(L, a);
    else fastq_split_kernel<OffT, false><<<(unsigned)blocks, SP_THREADS, dyn, st>>>(L, a);
    return cudaGetLastError();
}
cudaError_t fastq_split_launch(const uint8_t* buf, int64_t begin, int64_t n, const void* line_end, bool wide, int64_t n_rows, uint32_t mask,
                               int64_t* off, uint8_t* desc_valid, uint8_t* const* out, const int64_t* cap, void* scratch, const ScanResult* scan,
                               int map_mode, unsigned long long* bad, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    return wide ? split_launch_t<uint64_t>(buf, begin, n, line_end, n_rows, mask, off, desc_valid, out, cap, scratch, scan, map_mode, bad, st)
                : split_launch_t<uint32_t>(buf, begin, n, line_end, n_rows, mask, off, desc_valid, out, cap, scratch, scan, map_mode, bad, st);
}

static int row_blocks(int64_t n_rows, int rows_per_block) {
    int64_t b = (n_rows + rows_per_block - 1) / rows_per_block;
    const int64_t cap = 148 * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename OffT>
static cudaError_t fields_launch_t(const uint8_t* buf, int64_t begin, int64_t n, const void* line_end, const int64_t* sel, int64_t n_rows,
                                   uint32_t* lens, uint8_t* desc_valid, int64_t* starts, const ScanResult* scan, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    FqLines<OffT> L{buf, reinterpret_cast<const OffT*>(line_end), begin, n, true};
    fastq_fields_kernel<OffT><<<row_blocks(n_rows, 256), 256, 0, st>>>(L, sel, n_rows, lens, desc_valid, starts, scan);
    return cudaGetLastError();
}
cudaError_t fastq_fields_launch(const uint8_t* buf, int64_t begin, int64_t n, const void* line_end, bool wide, const int64_t* sel,
                                int64_t n_rows, uint32_t* lens, uint8_t* desc_valid, int64_t* starts, const ScanResult* scan, cudaStream_t st) {
    return wide ? fields_launch_t<uint64_t>(buf, begin, n, line_end, sel, n_rows, lens, desc_valid, starts, scan, st)
                : fields_launch_t<uint32_t>(buf, begin, n, line_end, sel, n_rows, lens, desc_valid, starts, scan, st);
}
template <typename OffT>
static cudaError_t gather_launch_t(const uint8_t* buf, int64_t begin, int64_t n, const void* line_end, const int64_t* sel, int64_t n_rows,
                                   int col, const uint32_t* lens, const int64_t* off, uint8_t* out, cudaStream_t st, int mode,
                                   unsigned long long* bad) {
    SrcFastq<OffT> fn{FqLines<OffT>{buf, reinterpret_cast<const OffT*>(line_end), begin, n, true}, sel, lens, col};
    return gather_span_launch(buf, fn, off, n_rows, n - begin, out, st, mode, bad);  // a column is never larger than the input
}
cudaError_t fastq_gather_launch(const uint8_t* buf, int64_t begin, int64_t n, const void* line_end, bool wide, const int64_t* sel,
                                int64_t n_rows, int col, const uint32_t* lens, const int64_t* off, uint8_t* out, cudaStream_t st, int mode,
                                unsigned long long* bad) {
    return wide ? gather_launch_t<uint64_t>(buf, begin, n, line_end, sel, n_rows, col, lens, off, out, st, mode, bad)
                : gather_launch_t<uint32_t>(buf, begin, n, line_end, sel, n_rows, col, lens, off, out, st, mode, bad);
}
cudaError_t gather_ranges_launch(const uint8_t* buf, const int64_t* start, const uint32_t* len, const int64_t* off, int64_t n_rows,
                                 int64_t total_bound, uint8_t* out, cudaStream_t st) {
    (void)len;
    return gather_span_launch(buf, SrcRanges{start}, off, n_rows, total_bound, out, st);
}

// the same through one of the reference's byte maps (EXB_MAP_*): *bad = ~0, or (position in out << 8) | byte, smallest position
cudaError_t gather_ranges_map_launch(const uint8_t* buf, const int64_t* start, const int64_t* off, int64_t n_rows, int64_t total_bound,
                                     uint8_t* out, int mode, unsigned long long* bad, cudaStream_t st) {
    return gather_span_launch(buf, SrcRanges{start}, off, n_rows, total_bound, out, st, mode, bad);
}

cudaError_t fastq_filter_launch(const uint32_t* seq_len, const uint32_t* gc, const uint32_t* qual_len, const int32_t* qsum, int64_t n,
                                const exb_predicate* preds, int n_preds, uint8_t* pass, int64_t* agg, const void* scan_ws,
                                cudaStream_t st) {
    FilterArgs a;
    a.scan = reinterpret_cast<const ScanResult*>(scan_ws);
    a.seq_len = seq_len;
    a.gc = gc;
    a.qual_len = qual_len;
    a.qsum = qsum;
    a.n = n;
    a.n_preds = n_preds;
    for (int i = 0; i < n_preds; i++) a.preds[i] = preds[i];
    a.pass = pass;
    a.agg = reinterpret_cast<long long*>(agg);
    cudaError_t e = cudaMemsetAsync(agg, 0, 8 * sizeof(int64_t), st);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    fastq_filter_kernel<<<row_blocks(n, 256 * 4), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ================================================================= FASTA headers
__device__ __forceinline__ bool is_ascii_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r'; }
__device__ __forceinline__ bool is_trim_ws(uint8_t c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); }

// Definition::from_str (noodles-fasta 0.27): name = up to the first ASCII
// whitespace; description = remainder trimmed; absent remainder = NULL.
__global__ void __launch_bounds__(256) fasta_headers_kernel(const uint8_t* __restrict__ buf, const int64_t* __restrict__ hdr_start,
                                                            const int64_t* __restrict__ hdr_end, int64_t n_rows, int64_t n_total,
                                                            uint32_t* __restrict__ lens, int64_t* __restrict__ desc_start,
                                                            uint8_t* __restrict__ desc_valid, unsigned long long* err_pos) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = hdr_start[i] + 1, e = hdr_end[i];
        if (e > s && e < n_total && buf[e - 1] == '\r') e--;  // CRLF; an unterminated last line keeps its CR
        int64_t p = s;
        while (p < e && !is_ascii_ws(buf[p])) p++;
        if (p == s) atomicMin(err_pos, (unsigned long long)hdr_start[i]);  // MissingName
        lens[i] = (uint32_t)(p - s);
        desc_start[i] = s;  // starts[0][i]: the id
        desc_start += n_rows;
        if (p < e) {
            int64_t ds = p + 1, de = e;
            while (ds < de && is_trim_ws(buf[ds])) ds++;
            while (de > ds && is_trim_ws(buf[de - 1])) de--;
            lens[n_rows + i] = (uint32_t)(de - ds);
            desc_start[i] = ds;
            desc_valid[i] = 1;
        } else {
            lens[n_rows + i] = 0;
            desc_start[i] = e;
            desc_valid[i] = 0;
        }
        desc_start -= n_rows;
    }
}
cudaError_t fasta_headers_launch(const uint8_t* buf, const int64_t* hdr_start, const int64_t* hdr_end, int64_t n_rows, int64_t n_total,
                                 uint32_t* lens, int64_t* desc_start, uint8_t* desc_valid, unsigned long long* err_pos,
                                 cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    fasta_headers_kernel<<<row_blocks(n_rows, 256), 256, 0, st>>>(buf, hdr_start, hdr_end, n_rows, n_total, lens, desc_start,
                                                                  desc_valid, err_pos);
    return cudaGetLastError();
}

// ================================================================= gc_content
__global__ void gc_from_prefix_kernel(const int64_t* __restrict__ seq_off, const int64_t* __restrict__ gc_prefix, int64_t n,
                                      float* __restrict__ out) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        // the reference counts in C ints; lengths beyond INT_MAX are outside what it can represent,
        // so the int64 -> float conversions below are the natural extension (round to nearest)
        long long len = seq_off[r + 1] - seq_off[r], gc = gc_prefix[r + 1] - gc_prefix[r];
        out[r] = len == 0 ? 0.0f : __fdiv_rn(__ll2float_rn(gc), __ll2float_rn(len));
    }
}
__global__ void gc_from_counts_kernel(const uint32_t* __restrict__ seq_len, const uint32_t* __restrict__ gc, int64_t n,
                                      float* __restrict__ out) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        out[r] = gc_fraction(gc[r], seq_len[r]);
}
cudaError_t gc_from_prefix_launch(const int64_t* seq_off, const int64_t* gc_prefix, int64_t n, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    gc_from_prefix_kernel<<<row_blocks(n, 256), 256, 0, st>>>(seq_off, gc_prefix, n, out);
    return cudaGetLastError();
}
cudaError_t gc_from_counts_launch(const uint32_t* seq_len, const uint32_t* gc, int64_t n, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    gc_from_counts_kernel<<<row_blocks(n, 256), 256, 0, st>>>(seq_len, gc, n, out);
    return cudaGetLastError();
}

// gc_content over a string column
// G/C bytes among data[s, e): aligned 8-byte words, the two ragged ends masked
__device__ __forceinline__ long long gc_count_range(const uint8_t* __restrict__ data, int64_t s, int64_t e, int64_t step_first, int64_t step) {
    // words w with index i = step_first, step_first + step, ... (so a warp can share one long row)
    const uint64_t* __restrict__ base = reinterpret_cast<const uint64_t*>((uintptr_t)(data + s) & ~(uintptr_t)7);
    const int64_t lead = (int64_t)((uintptr_t)(data + s) & 7);
    const int64_t nbytes = lead + (e - s);
    const int64_t nwords = (nbytes + 7) >> 3;
    long long cnt = 0;
    for (int64_t i = step_first; i < nwords; i += step) {
        uint64_t w = base[i];
        if (i == 0) w &= ~0ull << (8 * lead);  // a zeroed byte is not G/C
        if (i == nwords - 1 && (nbytes & 7)) w &= ~0ull >> (8 * (8 - (nbytes & 7)));
        cnt += __popc(gc_bytes((uint32_t)w)) + __popc(gc_bytes((uint32_t)(w >> 32)));
    }
    return cnt;
}

// the same count for one thread's own row, in 16-byte aligned chunks (half as many load instructions as 8-byte words)
__device__ __forceinline__ long long gc_count_row16(const uint8_t* __restrict__ data, int64_t s, int64_t e) {
    const uintptr_t a0 = (uintptr_t)(data + s) & ~(uintptr_t)15;
    const int lead = (int)((uintptr_t)(data + s) & 15);
    const int64_t nbytes = lead + (e - s);
    const int64_t nch = (nbytes + 15) >> 4;
    const uint4* __restrict__ base = reinterpret_cast<const uint4*>(a0);
    long long cnt = 0;
    for (int64_t i = 0; i < nch; i++) {
        const uint4 v = base[i];
        uint32_t m = (gc_bytes(v.x) >> 7) | (gc_bytes(v.y) >> 6) | (gc_bytes(v.z) >> 5) | (gc_bytes(v.w) >> 4);  // 16 flag bits, any order
        // bit layout: word k's byte b sits at bit 8 b + k; mask the bytes outside [lead, nbytes) of the first / last chunk
        if (i == 0 && lead) {
            uint32_t keep = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (4 * k + b >= lead) keep |= 1u << (8 * b + k);
            m &= keep;
        }
        if (i == nch - 1 && (nbytes & 15)) {
            const int last = (int)(nbytes & 15);
            uint32_t keep = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (4 * k + b < last) keep |= 1u << (8 * b + k);
            m &= keep;
        }
        cnt += __popc(m);
    }
    return cnt;
}

// One THREAD per row for rows up to GC_THREAD_ROW bytes (a warp's 32 rows are adjacent in the column, so its loads
// stay inside a few cache lines); longer rows are handed to the whole warp afterwards.
constexpr int64_t GC_THREAD_ROW = 1024;
__global__ void __launch_bounds__(256) gc_content_kernel(const int64_t* __restrict__ off, const uint8_t* __restrict__ data, int64_t n_rows,
                                                         float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r0 = tid - lane; r0 < n_rows; r0 += nthreads) {  // warp-uniform loop: r0 = the warp's first row
        const int64_t r = r0 + lane;
        int64_t s = 0, e = 0;
        if (r < n_rows) {
            s = off[r];
            e = off[r + 1];
        }
        const int64_t len = e - s;
        const bool is_long = len > GC_THREAD_ROW;
        if (r < n_rows && !is_long) {
            const long long cnt = len > 0 ? gc_count_row16(data, s, e) : 0;
            out[r] = len == 0 ? 0.0f : __fdiv_rn(__ll2float_rn(cnt), __ll2float_rn((long long)len));
        }
        uint32_t longs = __ballot_sync(0xffffffffu, is_long);
        while (longs) {
            const int l = __ffs((int)longs) - 1;
            longs &= longs - 1;
            const int64_t ls = __shfl_sync(0xffffffffu, s, l), le = __shfl_sync(0xffffffffu, e, l);
            long long cnt = gc_count_range(data, ls, le, lane, 32);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            if (lane == 0) out[r0 + l] = __fdiv_rn(__ll2float_rn(cnt), __ll2float_rn((long long)(le - ls)));
        }
    }
}
cudaError_t gc_content_launch(const int64_t* off, const uint8_t* data, int64_t n_rows, float* out, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    gc_content_kernel<<<row_blocks(n_rows, 256), 256, 0, st>>>(off, data, n_rows, out);
    return cudaGetLastError();
}

// ================================================================= LUT map / Phred decode
// 256-entry table in shared memory; entry 0 = invalid.  16 bytes per thread per step.
// the reference's per-byte tables (0 = throws): reverse_complement (module.cpp:30-69), complement (:81-121),
// transcribe (:212-249), reverse_transcribe (:168-203)
__device__ __forceinline__ void fill_map_lut(uint8_t* lut, int mode) {
    if (mode == EXB_MAP_REVERSE_COMPLEMENT) {
        lut['A'] = 'C'; lut['T'] = 'G'; lut['C'] = 'A'; lut['G'] = 'T';
    } else if (mode == EXB_MAP_COMPLEMENT) {
        lut['A'] = 'T'; lut['T'] = 'A'; lut['C'] = 'G'; lut['G'] = 'C';
    } else if (mode == EXB_MAP_TRANSCRIBE) {
        lut['A'] = 'A'; lut['C'] = 'C'; lut['G'] = 'G'; lut['T'] = 'U';
    } else {
        lut['A'] = 'A'; lut['C'] = 'C'; lut['G'] = 'G'; lut['U'] = 'T';
    }
}

__global__ void __launch_bounds__(256) seq_map_kernel(const uint8_t* __restrict__ in, int64_t n, int mode, uint8_t* __restrict__ out,
                                                      unsigned long long* bad_pos) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) fill_map_lut(lut, mode);
    __syncthreads();
    const bool aligned = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const int64_t nvec = aligned ? (n >> 4) : 0;
    const uint4* vin = reinterpret_cast<const uint4*>(in);
    uint4* vout = reinterpret_cast<uint4*>(out);
    unsigned long long bad = ~0ull;
    auto map4 = [&](uint32_t x, int64_t pos) -> uint32_t {
        uint32_t b0 = lut[x & 0xFF], b1 = lut[(x >> 8) & 0xFF], b2 = lut[(x >> 16) & 0xFF], b3 = lut[x >> 24];
        if (b0 == 0 || b1 == 0 || b2 == 0 || b3 == 0) {
            int k = b0 == 0 ? 0 : (b1 == 0 ? 1 : (b2 == 0 ? 2 : 3));
            unsigned long long p = (unsigned long long)(pos + k);
            bad = p < bad ? p : bad;
        }
        return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 x = vin[i], y;
        y.x = map4(x.x, i * 16);
        y.y = map4(x.y, i * 16 + 4);
        y.z = map4(x.z, i * 16 + 8);
        y.w = map4(x.w, i * 16 + 12);
        vout[i] = y;
    }
    for (int64_t i = nvec * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t b = lut[in[i]];
        if (b == 0) bad = (unsigned long long)i < bad ? (unsigned long long)i : bad;
        out[i] = b;
    }
    if (bad != ~0ull) atomicMin(bad_pos, bad);
}
cudaError_t seq_map_launch(const uint8_t* in, int64_t n, int mode, uint8_t* out, unsigned long long* bad_pos, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(bad_pos, 0xFF, 8, st);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    int64_t b = ((n >> 4) + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    seq_map_kernel<<<(int)b, 256, 0, st>>>(in, n, mode, out, bad_pos);
    return cudaGetLastError();
}

__device__ __forceinline__ void fill_map_lut_fwd(uint8_t* lut, int mode) { fill_map_lut(lut, mode); }

// ================================================================= translate_dna_to_aa (module.cpp:260-360)
// Rows whose length is a multiple of 3 tile the byte column in whole codons, so as long as every earlier row is well
// formed "protein of the concatenation" = "concatenation of the proteins": one thread per amino acid, three bytes in,
// one byte out through a 64-entry table indexed 16 a + 4 b + c (A C G T = 0 1 2 3).  Errors follow the reference's
// order (rows in order, a row's length before its codons): kernel 1 finds the first row with a bad length, kernel 2
// only looks at codons that lie before that row.
constexpr long long TR_NONE = 0x7F7F7F7F7F7F7F7Fll;  // what cudaMemsetAsync(0x7F) leaves: larger than any row or offset
__global__ void __launch_bounds__(256) translate_check_kernel(const int64_t* __restrict__ off, int64_t n_rows, long long* __restrict__ status) {
    long long bad = TR_NONE;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        if ((off[r + 1] - off[r]) % 3 != 0 && r < bad) bad = r;
    if (bad != TR_NONE) atomicMin(&status[0], bad);
}
__device__ __forceinline__ int base_code(uint32_t c) { return c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : -1))); }
__global__ void __launch_bounds__(256) translate_kernel(const int64_t* __restrict__ off, const uint8_t* __restrict__ data, int64_t n_rows,
                                                        uint8_t* __restrict__ out, long long* __restrict__ status) {
    __shared__ uint8_t table[64];
    if (threadIdx.x < 64) table[threadIdx.x] = (uint8_t)"KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"[threadIdx.x];
    __syncthreads();
    const int64_t base = off[0];
    const long long bad_row = status[0];  // written by the kernel before this one on the same stream
    const int64_t end = bad_row == TR_NONE ? off[n_rows] : off[bad_row];
    const int64_t n_aa = (end - base) / 3;
    long long bad = TR_NONE;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_aa; j += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t* p = data + base + 3 * j;
        const int a = base_code(p[0]), b = base_code(p[1]), c = base_code(p[2]);
        if ((a | b | c) < 0) {
            const long long pos = base + 3 * j;
            bad = pos < bad ? pos : bad;
            out[j] = 0;
        } else {
            out[j] = table[16 * a + 4 * b + c];
        }
    }
    if (bad != TR_NONE) atomicMin(&status[1], bad);
}
__global__ void translate_finish_kernel(long long* status) {
    if (threadIdx.x < 2 && status[threadIdx.x] == TR_NONE) status[threadIdx.x] = -1;
}
cudaError_t translate_launch(const int64_t* off, const uint8_t* data, int64_t n_rows, int64_t n_bytes_bound, uint8_t* out, long long* status,
                             cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(status, 0x7F, 16, st);  // both words = TR_NONE
    if (e != cudaSuccess) return e;
    if (n_rows > 0) {
        translate_check_kernel<<<row_blocks(n_rows, 256), 256, 0, st>>>(off, n_rows, status);
        const int64_t n_aa = n_bytes_bound / 3 + 1;
        const int64_t b = (n_aa + 255) / 256;
        translate_kernel<<<(unsigned)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16), 256, 0, st>>>(off, data, n_rows, out, status);
    }
    translate_finish_kernel<<<1, 32, 0, st>>>(status);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) quality_decode_kernel(const uint8_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
    // 4 input bytes -> one 16-byte store per thread per step
    const bool aligned = ((((uintptr_t)in) & 3) | (((uintptr_t)out) & 15)) == 0;
    const int64_t nq = aligned ? (n >> 2) : 0;
    const uint32_t* win = reinterpret_cast<const uint32_t*>(in);
    int4* vout = reinterpret_cast<int4*>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t x = win[i];
        int4 y;
        y.x = (int)(signed char)(x & 0xFF) - 33;
        y.y = (int)(signed char)((x >> 8) & 0xFF) - 33;
        y.z = (int)(signed char)((x >> 16) & 0xFF) - 33;
        y.w = (int)(signed char)(x >> 24) - 33;
        vout[i] = y;
    }
    for (int64_t i = nq * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int)(signed char)in[i] - 33;
}
cudaError_t quality_decode_launch(const uint8_t* in, int64_t n, int32_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int64_t b = ((n >> 2) + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    quality_decode_kernel<<<(int)b, 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

}  // namespace exb
