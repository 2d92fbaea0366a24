// fastq_scan.cu -- single-pass FASTQ line/record scan for sm_100a (v5: warp tiles).
//
// Replaces the record loop of noodles-fastq 0.8 Reader::read_record as driven
// by exon 0.2.6's FASTQ batch reader (call sites: rust/src/arrow_reader.rs:
// 104-118,125-153 in the reference).  Instead of reading a
// record at a time it reads every input byte exactly once.
//
// The unit of work is a WARP TILE: 2 KiB of input, owned by one warp from the
// global load to the last store.  There is no block-level barrier anywhere in
// the steady state; a persistent grid of 148 x (CTAs per SM) CTAs pulls tile
// ids from an atomic ticket, so the 40 warps of an SM are 40 independent
// pipelines whose stalls (HBM latency, the chain look-back) overlap freely.
// Each warp double-buffers: the cp.async (LDGSTS, L2-only) of tile k+1 is in
// flight while tile k is analysed, so ~80 KiB per SM are outstanding at all
// times -- what Little's law asks for at 6.5 TB/s.
//
// Per tile (lane l owns the 64 contiguous bytes [64l, 64l+64)):
//  A. analysis, byte-parallel and branch-free: 64-bit newline mask, signed
//     byte sum, G/C mask; one packed warp scan gives every lane its
//     tile-level prefix (newlines | byte sum) [+ G/C].  The tile's newline
//     count is published at once (chain word, status in the top bits).
//     Per-16-byte-chunk byte-sum prefixes go to shared memory (one STS.128),
//     newline positions are scattered, in order, into an event list.
//  B. tail: what follows the tile's last newline (start, byte sum, G/C) is
//     packed into ONE 64-bit word and published -- it depends on no other tile.
//  C. events, one newline per lane: prefix byte sum / G,C count at the newline
//     (chunk prefix + a 0/1-weight IDP.4A over the chunk's head).
//  D. chaining: decoupled look-back of the newline COUNT only (32 predecessors
//     per round); the line that is open at the tile's first byte is resolved
//     from the predecessors' tail words.
//  E. emission: line length, G/C count (sequence lines) and Phred sum (quality
//     lines) are differences of the prefixes at consecutive newlines; FASTQ's
//     strict 4-line phase (global line index mod 4) disambiguates '@' / '+'
//     inside quality strings.  Optionally the per-record predicate is applied
//     here and only COUNT / sums leave the kernel (fused filter).
//
// Outputs are single-writer stores (no atomics except the rare error path and
// one aggregate flush per warp):
//   line_end[g]            position of the newline ending line g      (F_LINES)
//   seq_len[r], gc[r]      per record                                 (F_SEQ)
//   qual_len[r], qsum[r]   per record; qsum = sum((signed char)c - 33) (F_QUAL)
// HBM traffic: input once + 16 B of chain state per 2 KiB + 4..16 B per line.
#include "common.cuh"
#include "exon_b200_internal.h"
#include "x87div.h"

namespace exb {

constexpr int WT_BYTES = 2048;            // bytes per warp tile
constexpr int WT_CHUNKS = WT_BYTES / 16;  // 128
constexpr int FQ_WARPS = 8;               // warps per CTA (independent of each other)
constexpr int FQ_THREADS = FQ_WARPS * 32;
constexpr int EV_CAP = 128;               // newline positions held at once (4 passes of 32)

// ---------------------------------------------------------------- chain / tail words
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

// tail word: [63:62] state (0 unpublished, 1 no newline in the tile, 2 has one)
//            [47:36] offset in the tile of the byte after its last newline (0..2048)
//            [35:24] G/C count of the part after it          (0..2048)
//            [23:0]  signed byte sum of that part            (|.| <= 2^18)
// Every field is self-contained in the word, so a relaxed 64-bit store publishes it.
__device__ __forceinline__ uint64_t tail_pack(uint32_t state, uint32_t rel, uint32_t g, int s) {
    return ((uint64_t)state << 62) | ((uint64_t)rel << 36) | ((uint64_t)g << 24) | (uint64_t)((uint32_t)s & 0xFFFFFFu);
}
__device__ __forceinline__ int tail_s_of(uint64_t w) { return ((int)((uint32_t)w << 8)) >> 8; }
__device__ __forceinline__ uint32_t tail_g_of(uint64_t w) { return (uint32_t)(w >> 24) & 0xFFFu; }
__device__ __forceinline__ uint32_t tail_rel_of(uint64_t w) { return (uint32_t)(w >> 36) & 0xFFFu; }

// ---------------------------------------------------------------- byte classification
// 16-bit equality mask of a 16-byte chunk, bits in byte order.  The 0x80 flags of two
// words are folded into one byte by IDP.4A with weights 1,2,4,8 / 16,32,64,128 (the
// products carry a factor 128 that one shift removes).
__device__ __forceinline__ uint32_t flags_to_mask16(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    uint32_t lo = __dp4a(m0, 0x08040201u, 0u);
    lo = __dp4a(m1, 0x80402010u, lo);
    uint32_t hi = __dp4a(m2, 0x08040201u, 0u);
    hi = __dp4a(m3, 0x80402010u, hi);
    return (lo >> 7) | (hi << 1);
}
// 0x80 per byte equal to the byte replicated in `pat` (< 0x80).  c7f = 0x7F7F7F7F and pat
// arrive in REGISTERS (the caller launders them through the kernel arguments) so that
// (x & c7f) ^ pat is one LOP3 instead of two immediate-form ones.
__device__ __forceinline__ uint32_t eq_flags(uint32_t x, uint32_t c7f, uint32_t pat) {
    const uint32_t t = (x & c7f) ^ pat;
    return ~((t + c7f) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t nl_mask16r(const uint4& v, uint32_t c7f, uint32_t pat) {
    return flags_to_mask16(eq_flags(v.x, c7f, pat), eq_flags(v.y, c7f, pat), eq_flags(v.z, c7f, pat), eq_flags(v.w, c7f, pat));
}
// 'G' (0x47) and 'C' (0x43) differ only in bit 2
__device__ __forceinline__ uint32_t gc_flags(uint32_t x, uint32_t c7b, uint32_t c7f, uint32_t pat) {
    const uint32_t t = (x & c7b) ^ pat;
    return ~((t + c7f) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t gc_mask16r(const uint4& v, uint32_t c7b, uint32_t c7f, uint32_t pat) {
    return flags_to_mask16(gc_flags(v.x, c7b, c7f, pat), gc_flags(v.y, c7b, c7f, pat), gc_flags(v.z, c7b, c7f, pat),
                           gc_flags(v.w, c7b, c7f, pat));
}

// ---------------------------------------------------------------- shared memory of one warp
template <int FLAGS>
struct FqWarpSmem {
    static constexpr bool kSeq = (FLAGS & EXB_F_SEQ) != 0, kQual = (FLAGS & EXB_F_QUAL) != 0;
    static constexpr int off_data = 0;                                   // 2 x 2 KiB (double buffer)
    static constexpr int off_cpre = off_data + 2 * WT_BYTES;             // int[128]: byte-sum prefix at each chunk
    static constexpr int off_gm = off_cpre + (kQual ? WT_CHUNKS * 4 : 0);  // u64[32]: G/C mask of each lane's run
    static constexpr int off_gex = off_gm + (kSeq ? 32 * 8 : 0);         // int[32]: G/C prefix at each lane's run
    static constexpr int off_ev = off_gex + (kSeq ? 32 * 4 : 0);         // u16[EV_CAP]: newline positions
    static constexpr int total = off_ev + EV_CAP * 2;
};

// tile-local byte index -> byte offset in the (chunk-swizzled) tile buffer
__device__ __forceinline__ int sidx(int li) {
    const int q = li >> 4;
    return ((q ^ ((q >> 3) & 3)) << 4) | (li & 15);
}

#ifdef EXB_FQ_TRACE
__device__ unsigned long long* g_fq_trace;  // 4 words per tile: t(publish), t(resolved), spins | last<<32, t(ticket)
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FQ_TRACE(slot, val) do { if (lane == 0 && g_fq_trace) g_fq_trace[tile * 4 + (slot)] = (val); } while (0)
#else
#define FQ_TRACE(slot, val) do { } while (0)
#endif

struct OpenLine {  // the line that is open at a tile's first byte
    int64_t start;
    int s, g;      // byte sum / G,C count of its part before the tile (mod 2^32)
};

// Walk the predecessors' tail words back to the tile that holds the open line's start
// (tile-1 unless lines are longer than a tile).  All 32 lanes call it; 32 tiles per round.
__device__ __forceinline__ OpenLine open_line_before(const uint64_t* tails, int64_t tile, int64_t origin, const FastqScanArgs& a) {
    const int lane = threadIdx.x & 31;
    OpenLine o;
    o.s = 0;
    o.g = 0;
    int64_t base = tile - 1;
    while (true) {
        const int64_t idx = base - lane;
        uint64_t w;
        if (idx >= 0) w = ld_relaxed_u64(tails + idx);
        else w = 2ull << 62;  // the state before tile 0 terminates the walk
        const uint32_t st = (uint32_t)(w >> 62);
        const uint32_t has = __ballot_sync(0xffffffffu, st == 2), emp = __ballot_sync(0xffffffffu, st == 0);
        const int first = has ? __ffs(has) - 1 : 32;
        const uint32_t need = first >= 32 ? 0xffffffffu : ((2u << first) - 1u);  // lanes up to and including `first`
        if (emp & need) continue;  // a needed predecessor has not published yet
        const bool take = lane <= first && idx >= 0;
        o.s += (int)__reduce_add_sync(0xffffffffu, take ? tail_s_of(w) : 0);
        o.g += (int)__reduce_add_sync(0xffffffffu, take ? tail_g_of(w) : 0u);
        if (has) {
            const int64_t fidx = base - first;
            const uint32_t rel = __shfl_sync(0xffffffffu, tail_rel_of(w), first);
            if (fidx >= 0) {
                o.start = origin + fidx * WT_BYTES + rel;
            } else if (a.prev) {  // the line began in an earlier range of a chained scan
                o.start = a.prev->open_line_start;
                o.s += (int)a.prev->tail_s;
                o.g += (int)a.prev->tail_g;
            } else {
                o.start = a.begin;
            }
            return o;
        }
        base -= 32;
    }
}

// ---------------------------------------------------------------- two-level chain of the newline counts
// A decoupled look-back advances its frontier by (window x tile bytes) per L2 round trip.
// With 2 KiB warp tiles that is far too slow for a flat chain (32 x 2 KiB per ~0.6 us =
// 0.1 TB/s), so the counts are chained at two levels:
//   level 1  cnt1[t]   = 0x80000000 | newlines of warp tile t            (plain store)
//   level 2  super[j]  = accumulator of the SUPER_TILES = 32 warp tiles [32j, 32j+32):
//                        (arrivals << 32 | sum), built with one atomicAdd per warp tile.
//                        The warp whose atomicAdd completes the group runs the look-back
//                        over super words (128 per round = 8 MiB of input per round trip)
//                        and overwrites the word with SUP_INC | inclusive line count.
// A warp tile's exclusive prefix = inclusive(super j-1) + sum of cnt1 of its earlier siblings.
constexpr int SUPER_TILES = 32;
constexpr uint64_t SUP_INC = 1ull << 63, SUP_VAL = (1ull << 62) - 1ull;
constexpr uint32_t CNT1_VALID = 0x80000000u;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Run by the warp that completed super tile j (j > 0): lines before the super tile.
__device__ __forceinline__ uint64_t super_lookback(const uint64_t* sup, int64_t j, uint64_t init) {
    const int lane = threadIdx.x & 31;
    uint64_t acc = 0;
    int64_t base = j - 1;
    while (true) {
        uint64_t w[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int64_t idx = base - (32 * m + lane);
            w[m] = idx >= 0 ? ld_relaxed_u64(sup + idx) : (SUP_INC | (idx == -1 ? (init & SUP_VAL) : 0ull));
        }
        bool done = false, stalled = false;
#pragma unroll
        for (int m = 0; m < 4; m++) {
            if (done || stalled) continue;
            const bool is_inc = (w[m] >> 63) != 0;
            const bool complete = !is_inc && (uint32_t)(w[m] >> 32) == (uint32_t)SUPER_TILES;  // every earlier super tile is full
            const uint32_t inc = __ballot_sync(0xffffffffu, is_inc), emp = __ballot_sync(0xffffffffu, !is_inc && !complete);
            const int first = inc ? __ffs(inc) - 1 : 32;
            const uint32_t need = first >= 32 ? 0xffffffffu : ((1u << first) - 1u);
            if (emp & need) {  // a needed group is still counting: poll again from this window
                stalled = true;
                base -= 32 * m;
                continue;
            }
            acc += __reduce_add_sync(0xffffffffu, lane < first ? (uint32_t)w[m] : 0u);
            if (inc) {
                acc += __shfl_sync(0xffffffffu, w[m] & SUP_VAL, first);
                done = true;
            }
        }
        if (done) return acc;
        if (!stalled) base -= 128;
    }
}

template <typename OffT, int FLAGS>
__global__ void __launch_bounds__(FQ_THREADS, 5) fastq_scan_kernel(const FastqScanArgs a, const uint32_t c7f, const uint32_t c7b) {
    using SM = FqWarpSmem<FLAGS>;
    constexpr bool kLines = (FLAGS & EXB_F_LINES) != 0;
    constexpr bool kSeq = SM::kSeq, kQual = SM::kQual;
    constexpr bool kFused = (FLAGS & EXB_F_FUSED) != 0;

    extern __shared__ __align__(16) uint8_t smem_all[];
    __shared__ uint4 s_wlut[17];  // s_wlut[k]: 0x01 in the first k bytes -- IDP.4A weights of a chunk's head

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 17) {
        const int k = threadIdx.x;
        auto w = [k](int b) -> uint32_t {  // word holding bytes [b, b+4)
            const int c = k - b;
            return c <= 0 ? 0u : (c >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - c))));
        };
        s_wlut[k] = make_uint4(w(0), w(4), w(8), w(12));
    }
    __syncthreads();  // the only block-wide barrier of the kernel

    uint8_t* sm = smem_all + warp * SM::total;
    int* s_cpre = reinterpret_cast<int*>(sm + SM::off_cpre);
    uint64_t* s_gm = reinterpret_cast<uint64_t*>(sm + SM::off_gm);
    int* s_gex = reinterpret_cast<int*>(sm + SM::off_gex);
    uint16_t* ev_pos = reinterpret_cast<uint16_t*>(sm + SM::off_ev);

    const uint8_t* __restrict__ buf = a.buf;
    const int64_t origin = a.begin & ~(int64_t)15;
    const int64_t n_tiles = a.n_tiles;
    const int64_t n_super = (n_tiles + SUPER_TILES - 1) / SUPER_TILES;
    uint64_t* tails = reinterpret_cast<uint64_t*>(a.slots);
    uint64_t* sup = tails + n_tiles;
    uint32_t* cnt1 = reinterpret_cast<uint32_t*>(sup + n_super);
    const uint64_t init = a.prev ? a.prev->total_lines : 0ull;
    const uint32_t pat_nl = c7f & 0x0A0A0A0Au, pat_gc = c7f & 0x43434343u;  // derived from an argument: stay in registers

    // fused aggregates of this warp (flushed once at the end)
    long long f_cnt = 0, f_qs = 0, f_ql = 0;

    // ---- staging of one tile into buffer b (asynchronous)
    const uint32_t dst_lane = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)((lane ^ ((lane >> 3) & 3)) << 4);
    auto issue = [&](int64_t tile, int b) {
        if (tile < n_tiles) {
            const int64_t base = origin + tile * WT_BYTES;
            const uint32_t dst = dst_lane + (uint32_t)(b * WT_BYTES);
            if (base >= a.begin && base + WT_BYTES <= a.n) {  // interior tile: no clipping
                const uint8_t* src = buf + base + lane * 16;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + i * 512), "l"(src + i * 512) : "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int64_t g = base + (int64_t)(i * 32 + lane) * 16;
                    const int64_t rem = a.n - g;
                    const int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);  // bytes beyond n are zero-filled
                    const uint8_t* src = nb > 0 ? buf + g : buf;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + i * 512), "l"(src), "r"(nb) : "memory");
                }
            }
        }
        cp_async_commit();
    };
    unsigned long long ticket_raw = 0;
    auto take_ticket = [&]() {
        if (lane == 0) ticket_raw = atomicAdd(a.ticket, 1ull);
    };
    auto ticket_value = [&]() -> int64_t { return (int64_t)__shfl_sync(0xffffffffu, ticket_raw, 0); };

    take_ticket();
    int64_t cur = ticket_value();
    if (cur == 0 && lane == 0 && a.prev) {  // chained range: errors of earlier ranges stay visible in the last result
        if (a.prev->err_pos != 0ull) atomicMax(&a.result->err_pos, a.prev->err_pos);
        if (a.prev->overflow) a.result->overflow = 1;
    }
    issue(cur, 0);
    take_ticket();
    int b = 0;

    while (cur < n_tiles) {
        const int64_t nxt = ticket_value();
        issue(nxt, b ^ 1);
        take_ticket();  // for the tile after next; its latency hides behind this tile's work
        cp_async_wait<1>();
        __syncwarp();

        const int64_t tile = cur;
        const int64_t tile_base = origin + tile * WT_BYTES;
        FQ_TRACE(3, gtime());
        uint8_t* sbytes = sm + b * WT_BYTES;
        const uint4* d = reinterpret_cast<const uint4*>(sbytes);

        // Rare per-launch patches: bytes before `begin` in the first chunk are filler; an
        // unterminated last line gets a virtual '\n' at index n.
        const bool has_begin_pad = (tile == 0 && a.begin != origin);
        const bool has_eof = a.is_final && (a.n >= tile_base && a.n < tile_base + WT_BYTES);
        if (has_begin_pad || has_eof) {
            if (lane == 0) {
                if (has_begin_pad)
                    for (int64_t i = origin; i < a.begin; i++) sbytes[sidx((int)(i - origin))] = 0;
                if (has_eof) {
                    const bool open = a.n > a.begin ? buf[a.n - 1] != '\n' : (a.prev && a.prev->open_line_start < a.n);
                    if (open) sbytes[sidx((int)(a.n - tile_base))] = '\n';
                }
            }
            __syncwarp();
        }
        auto any_byte = [&](int64_t abs_pos) -> int {  // byte at an absolute offset < tile end that may lie before this tile
            const int64_t li = abs_pos - tile_base;
            if (li >= 0) return sbytes[sidx((int)li)];
            return (abs_pos >= (a.prev ? 0 : a.begin)) ? (int)buf[abs_pos] : -1;
        };

        // ---- A. analysis of the lane's 64-byte run
        uint64_t pm, gm = 0;
        unsigned long long sup_old = 0;  // lane 0: the super tile's accumulator before this tile joined
        int ex_s = 0, ex_g = 0, ex_cnt, n_events, total_s = 0, total_g = 0;
        {
            const int x = (lane >> 1) & 3, rb = 4 * lane;
            const uint4 c0 = d[rb + (0 ^ x)], c1 = d[rb + (1 ^ x)], c2 = d[rb + (2 ^ x)], c3 = d[rb + (3 ^ x)];
            pm = ((uint64_t)(nl_mask16r(c2, c7f, pat_nl) | (nl_mask16r(c3, c7f, pat_nl) << 16)) << 32) |
                 (nl_mask16r(c0, c7f, pat_nl) | (nl_mask16r(c1, c7f, pat_nl) << 16));
            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            if (kQual) {
                s0 = sbyte_sum16(c0, 0);
                s1 = sbyte_sum16(c1, s0);
                s2 = sbyte_sum16(c2, s1);
                s3 = sbyte_sum16(c3, s2);
            }
            int gtot = 0;
            if (kSeq) {
                gm = ((uint64_t)(gc_mask16r(c2, c7b, c7f, pat_gc) | (gc_mask16r(c3, c7b, c7f, pat_gc) << 16)) << 32) |
                     (gc_mask16r(c0, c7b, c7f, pat_gc) | (gc_mask16r(c1, c7b, c7f, pat_gc) << 16));
                gtot = __popcll(gm);
            }
            const int cnt = __popcll(pm);
            // newlines (<= 2048 per tile: 12 bits) and byte sum (|.| <= 2^18: 20 bits signed) share one scan word
            const uint32_t packed = ((uint32_t)cnt << 20) + (uint32_t)(kQual ? s3 : gtot);
            const uint32_t incl = warp_incl_scan_u32(packed);
            const uint32_t ex = incl - packed;
            const int ex_lo = ((int)(ex << 12)) >> 12;
            ex_cnt = (int)((ex - (uint32_t)ex_lo) >> 20);
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            const int tot_lo = ((int)(tot << 12)) >> 12;
            n_events = (int)((tot - (uint32_t)tot_lo) >> 20);
            if (kQual) {
                ex_s = ex_lo;
                total_s = tot_lo;
                if (kSeq) {
                    const uint32_t gi = warp_incl_scan_u32((uint32_t)gtot);
                    ex_g = (int)(gi - (uint32_t)gtot);
                    total_g = (int)__shfl_sync(0xffffffffu, gi, 31);
                }
            } else if (kSeq) {
                ex_g = ex_lo;
                total_g = tot_lo;
            }
            // the count is all the chain needs: publish it before anything else
            FQ_TRACE(0, gtime());
            if (lane == 0) {
                st_relaxed_u32(cnt1 + tile, CNT1_VALID | (uint32_t)n_events);
                sup_old = atomicAdd(reinterpret_cast<unsigned long long*>(sup + (tile >> 5)), (1ull << 32) | (unsigned long long)n_events);
            }
            if (kQual) *reinterpret_cast<int4*>(s_cpre + rb) = make_int4(ex_s, ex_s + s0, ex_s + s1, ex_s + s2);
            if (kSeq) {
                s_gm[lane] = gm;
                s_gex[lane] = ex_g;
            }
        }
        // newline positions with rank in [win_lo, win_lo + EV_CAP), in order
        auto scatter = [&](int win_lo) {
            int rank = ex_cnt - win_lo;
            const int p0 = lane * 64;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t m = h ? (uint32_t)(pm >> 32) : (uint32_t)pm;
                while (m) {
                    const int k = __ffs((int)m) - 1;
                    m &= m - 1;
                    if ((unsigned)rank < (unsigned)EV_CAP) ev_pos[rank] = (uint16_t)(p0 + h * 32 + k);
                    rank++;
                }
            }
        };
        scatter(0);
        __syncwarp();

        // prefix sums at a newline position (tile-local): byte sum / G,C count of everything before it
        auto event_prefix = [&](int pos, int& ps, int& pg) {
            if (kQual) {
                const int q = pos >> 4;
                const uint4 v = d[q ^ ((q >> 3) & 3)];
                const uint4 w = s_wlut[pos & 15];
                int acc = s_cpre[q];
                acc = __dp4a((int)v.x, (int)w.x, acc);
                acc = __dp4a((int)v.y, (int)w.y, acc);
                acc = __dp4a((int)v.z, (int)w.z, acc);
                acc = __dp4a((int)v.w, (int)w.w, acc);
                ps = acc;
            }
            if (kSeq) pg = s_gex[pos >> 6] + __popcll(s_gm[pos >> 6] & low_bits64(pos & 63));
        };

        // ---- B. tail word: position / sums after the tile's last newline (local information only)
        int last_pos = -1, last_ps = 0, last_pg = 0;
        if (n_events > 0) {
            const uint32_t have = __ballot_sync(0xffffffffu, pm != 0);
            const int top = 31 - __clz((int)have);
            const int hi = 63 - __clzll((long long)pm);  // meaningful in lane `top`
            last_pos = top * 64 + __shfl_sync(0xffffffffu, hi, top);
            event_prefix(last_pos, last_ps, last_pg);
            if (lane == 0)
                st_relaxed_u64(tails + tile, tail_pack(2, (uint32_t)(last_pos + 1), (uint32_t)(total_g - last_pg), total_s - (last_ps + 10)));
        } else if (lane == 0) {
            st_relaxed_u64(tails + tile, tail_pack(1, 0, (uint32_t)total_g, total_s));
        }

        // ---- C. prefixes of the first 32 events (they do not depend on other tiles)
        int pos = 0, ps = 0, pg = 0;
        if (lane < n_events) {
            pos = ev_pos[lane];
            event_prefix(pos, ps, pg);
        }

        // ---- D. chaining: global line index of the tile's first newline; the open line
        uint64_t excl;
        {
            const int64_t j = tile >> 5;
            const int i = (int)(tile & 31);
            const int group = (int)((n_tiles - j * SUPER_TILES) < SUPER_TILES ? (n_tiles - j * SUPER_TILES) : SUPER_TILES);
            const unsigned long long old = __shfl_sync(0xffffffffu, sup_old, 0);
            uint64_t sup_excl = init;
            bool have_sup = j == 0;
            if ((int)(old >> 32) == group - 1) {  // this warp completed the super tile: chain it
                if (j > 0) sup_excl = super_lookback(sup, j, init);
                have_sup = true;
                if (lane == 0) st_relaxed_u64(sup + j, SUP_INC | ((sup_excl + (old & 0xffffffffull) + (uint64_t)n_events) & SUP_VAL));
            }
            uint32_t c = 0;
            unsigned long long spins = 0;
            while (true) {
                spins++;
                if (lane < i && !(c & CNT1_VALID)) c = ld_relaxed_u32(cnt1 + j * SUPER_TILES + lane);
                uint64_t w = SUP_INC;
                if (!have_sup) w = ld_relaxed_u64(sup + j - 1);
                const bool ok1 = __all_sync(0xffffffffu, lane >= i || (c & CNT1_VALID));
                if (!have_sup && (w >> 63)) {
                    sup_excl = w & SUP_VAL;
                    have_sup = true;
                }
                if (ok1 && have_sup) break;
            }
            excl = sup_excl + __reduce_add_sync(0xffffffffu, lane < i ? (c & ~CNT1_VALID) : 0u);
            FQ_TRACE(1, gtime());
            FQ_TRACE(2, spins | ((unsigned long long)((int)(old >> 32) == group - 1) << 32));
        }
        OpenLine open;
        open.start = 0;
        open.s = open.g = 0;
        const bool is_last = tile == n_tiles - 1;
        if (n_events > 0 || is_last) open = open_line_before(tails, tile, origin, a);

        if (is_last && lane == 0) {  // final state of this range (chaining / host)
            a.result->total_lines = excl + (uint64_t)n_events;
            if (n_events > 0) {
                a.result->open_line_start = tile_base + last_pos + 1;
                a.result->tail_s = total_s - (last_ps + 10);
                a.result->tail_g = total_g - last_pg;
            } else {
                a.result->open_line_start = open.start;
                a.result->tail_s = (int64_t)open.s + total_s;
                a.result->tail_g = (int64_t)open.g + total_g;
            }
        }

        // ---- E. emission, 32 events per pass
        int c_pos = 0, c_ps = 0, c_pg = 0;  // last event of the previous pass (ps includes that newline)
        for (int lo = 0; lo < n_events; lo += 32) {
            if (lo > 0) {
                if ((lo & (EV_CAP - 1)) == 0) {  // more newlines than the event window holds (rare): refill it
                    __syncwarp();
                    scatter(lo);
                    __syncwarp();
                }
                if (lo + lane < n_events) {
                    pos = ev_pos[(lo & (EV_CAP - 1)) + lane];
                    event_prefix(pos, ps, pg);
                }
            }
            const bool active = lo + lane < n_events;
            int ppos = __shfl_up_sync(0xffffffffu, pos, 1), pps = 0, ppg = 0;
            if (kQual) pps = __shfl_up_sync(0xffffffffu, ps, 1) + 10;
            if (kSeq) ppg = __shfl_up_sync(0xffffffffu, pg, 1);
            if (lane == 0) {
                ppos = c_pos;
                pps = c_ps;
                ppg = c_pg;
            }
            if (active) {
                const uint64_t g = excl + (uint64_t)(lo + lane);  // global index of the line this newline ends
                const bool first_of_tile = (lo + lane) == 0;
                uint32_t len;
                int ssum = 0, gsum = 0, first_byte;
                int64_t line_start;
                if (!first_of_tile) {
                    len = (uint32_t)(pos - ppos - 1);
                    if (kQual) ssum = ps - pps;
                    if (kSeq) gsum = pg - ppg;
                    line_start = tile_base + ppos + 1;
                } else {  // the line that was open when the tile began
                    line_start = open.start;
                    len = (uint32_t)(tile_base - open.start) + (uint32_t)pos;
                    if (kQual) ssum = ps + open.s;
                    if (kSeq) gsum = pg + open.g;
                }
                const int64_t e = tile_base + pos;
                if (g < a.max_lines) {
                    // a CR directly before a real LF is stripped; the virtual '\n' at EOF strips nothing
                    uint32_t cr = 0;
                    if (len > 0 && !(a.is_final && e == a.n)) cr = (pos > 0 ? (int)sbytes[sidx(pos - 1)] : any_byte(e - 1)) == '\r';
                    len -= cr;
                    const int ph = (int)(g & 3);
                    const uint64_t r = g >> 2;
                    if (kLines) {
                        if (g < (uint64_t)a.line_cap) reinterpret_cast<OffT*>(a.line_end)[g] = (OffT)e;
                        else a.result->overflow = 1;
                    }
                    if ((ph & 1) == 0) {  // header / plus line: its first byte must be '@' / '+'
                        first_byte = first_of_tile ? any_byte(line_start) : (int)sbytes[sidx(ppos + 1)];
                        if (first_byte != (ph == 0 ? '@' : '+')) atomicMax(&a.result->err_pos, ~(unsigned long long)line_start);
                    } else if (ph == 1) {
                        if (kSeq) {
                            if (r < (uint64_t)a.rec_cap) {
                                a.seq_len[r] = len;
                                a.gc[r] = (uint32_t)gsum;
                            } else {
                                a.result->overflow = 1;
                            }
                        }
                    } else if (kQual) {
                        const int qs = ssum - 13 * (int)cr - 33 * (int)len;
                        if (kFused) {
                            bool ok = true;
                            for (int i = 0; i < a.n_fused; i++) {
                                const exb_predicate p = a.fused[i];
                                ok = ok && (p.field == EXB_P_MEAN_QUALITY ? exb_mean_cmp((int64_t)qs, len, p.op, p.value)
                                                                          : exb_cmp((double)len, p.op, p.value));
                            }
                            if (ok) {
                                f_cnt += 1;
                                f_qs += qs;
                                f_ql += len;
                            }
                        } else if (r < (uint64_t)a.rec_cap) {
                            a.qual_len[r] = len;
                            a.qsum[r] = qs;
                        } else {
                            a.result->overflow = 1;
                        }
                    }
                }
            }
            c_pos = __shfl_sync(0xffffffffu, pos, 31);
            if (kQual) c_ps = __shfl_sync(0xffffffffu, ps, 31) + 10;
            if (kSeq) c_pg = __shfl_sync(0xffffffffu, pg, 31);
        }

        __syncwarp();  // the event list / prefix arrays are rewritten by the next tile
        cur = nxt;
        b ^= 1;
    }
    cp_async_wait<0>();

    if (kFused) {  // one flush per warp
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) {
            f_cnt += __shfl_xor_sync(0xffffffffu, f_cnt, dd);
            f_qs += __shfl_xor_sync(0xffffffffu, f_qs, dd);
            f_ql += __shfl_xor_sync(0xffffffffu, f_ql, dd);
        }
        if (lane == 0 && f_cnt != 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(a.fused_agg) + 0, (unsigned long long)f_cnt);
            atomicAdd(reinterpret_cast<unsigned long long*>(a.fused_agg) + 3, (unsigned long long)f_qs);
            atomicAdd(reinterpret_cast<unsigned long long*>(a.fused_agg) + 4, (unsigned long long)f_ql);
        }
    }
}

// ------------------------------------------------------------------ launcher
int64_t fastq_scan_tiles(int64_t begin, int64_t n, int is_final) {
    const int64_t origin = begin & ~(int64_t)15;
    // final range: +1 byte of room for the virtual terminator at n
    const int64_t t = (n + (is_final ? 1 : 0) - origin + WT_BYTES - 1) / WT_BYTES;
    return t > 0 ? t : 1;
}
int64_t fastq_scan_chain_bytes(int64_t n_tiles) {  // tail word + count word per tile, one accumulator per 32 tiles
    return n_tiles * 12 + ((n_tiles + SUPER_TILES - 1) / SUPER_TILES) * 8 + 16;
}

template <typename OffT, int FLAGS>
static cudaError_t launch_one(const FastqScanArgs& a, cudaStream_t st) {
    constexpr int smem = FqWarpSmem<FLAGS>::total * FQ_WARPS;
    auto kern = fastq_scan_kernel<OffT, FLAGS>;
    static int ctas_per_sm = 0, n_sm = 0;  // per template instance; one device type per process
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int dev = 0, occ = 0, sms = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FQ_THREADS, smem)) != cudaSuccess) return e;
        n_sm = sms;
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    int64_t grid = (a.n_tiles + FQ_WARPS - 1) / FQ_WARPS;
    const int64_t persistent = (int64_t)n_sm * ctas_per_sm;
    if (grid > persistent) grid = persistent;
    // 0x7F7F7F7F / 0x7B7B7B7B travel as arguments so that ptxas keeps them in registers (see eq_flags)
    kern<<<dim3((unsigned)grid), dim3(FQ_THREADS), smem, st>>>(a, 0x7F7F7F7Fu, 0x7B7B7B7Bu);
    return cudaGetLastError();
}

template <typename OffT>
static cudaError_t launch_fastq(const FastqScanArgs& a, int flags, cudaStream_t st) {
    if (flags & EXB_F_FUSED) {  // quality-line predicates + COUNT, nothing per record leaves the kernel
        return (flags & EXB_F_LINES) ? launch_one<OffT, EXB_F_FUSED | EXB_F_QUAL | EXB_F_LINES>(a, st)
                                     : launch_one<OffT, EXB_F_FUSED | EXB_F_QUAL>(a, st);
    }
    switch (flags & 7) {
    case 0: return launch_one<OffT, 0>(a, st);
    case 1: return launch_one<OffT, 1>(a, st);
    case 2: return launch_one<OffT, 2>(a, st);
    case 3: return launch_one<OffT, 3>(a, st);
    case 4: return launch_one<OffT, 4>(a, st);
    case 5: return launch_one<OffT, 5>(a, st);
    case 6: return launch_one<OffT, 6>(a, st);
    default: return launch_one<OffT, 7>(a, st);
    }
}

#ifdef EXB_FQ_TRACE
extern "C" __attribute__((visibility("default"))) int exb_debug_set_fq_trace(void* d_trace) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(d_trace);
    return (int)cudaMemcpyToSymbol(g_fq_trace, &p, sizeof(p));
}
#endif

cudaError_t fastq_scan_launch(const FastqScanArgs& a, int flags, bool wide_offsets, cudaStream_t st) {
    return wide_offsets ? launch_fastq<uint64_t>(a, flags, st) : launch_fastq<uint32_t>(a, flags, st);
}

}  // namespace exb
