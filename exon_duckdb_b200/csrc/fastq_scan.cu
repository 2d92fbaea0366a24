// fastq_scan.cu -- FASTQ line/record scan for sm_100a (v7: TMA-fed chain-free tile pass + emit pass).
//
// Replaces the record loop of noodles-fastq 0.8 Reader::read_record as driven
// by exon 0.2.6's FASTQ batch reader (call sites: rust/src/arrow_reader.rs:
// 104-118,125-153 in the reference).  Instead of reading a
// record at a time it reads every input byte exactly once.
//
// Why two kernels.  What a FASTQ line MEANS depends on its global line index
// (strict 4-line records: index mod 4 picks header / sequence / plus / quality,
// index / 4 is the record).  A single-pass kernel therefore has to chain a
// prefix count across tiles; measured on B200 (profiles/r01c_*_trace.txt) the
// chain costs ~10 us per tile even with a two-level look-back, because every
// tile waits for the slowest earlier one, and any warp that waits while holding
// a prefetched tile makes that tile late for everybody else.  So the work is
// split where the dependency is:
//
//  K1  fastq_tile_kernel   (all the byte work, NO inter-tile dependency)
//      A warp owns a 4 KiB tile.  A persistent grid walks the tiles with a
//      static stride; each warp double-buffers its tiles with TMA: one elected
//      lane issues ONE cp.async.bulk.tensor (UTMALDG) per tile -- a 32 x 128 B
//      box of the file viewed as rows of 128 bytes, SWIZZLE_128B -- completing
//      on an mbarrier.  Lane l owns row l (128 contiguous bytes); the hardware
//      swizzle makes its eight LDS.128 bank-conflict free.  Per lane: two
//      64-bit newline masks, signed byte sums, G/C masks; warp scans.  Then,
//      one newline per lane, the prefix byte sum / G,C count at the newline
//      (per-chunk prefixes + a 0/1-weight IDP.4A over the chunk's head).
//      Per tile it writes: the newline count, one 8-byte record per newline and
//      one 64-bit tail word describing what follows the last newline.
//  --  exclusive scan of the per-tile counts (1.7 M words for 7 GB: microseconds)
//  K2  fastq_emit_kernel   (light: 8 bytes per LINE instead of ~90)
//      With the global line index known, line length, G/C count (sequence
//      lines) and Phred sum (quality lines) are differences of consecutive
//      records; the line that straddles a tile edge is completed from the
//      predecessors' tail words.  '@' / '+' are validated by line phase.
//
//  Fused COUNT flavour (exb_fastq_scan_filter): K1 evaluates the predicate for
//  every line as if it were a quality line and adds it to one of FOUR buckets
//  (one per possible phase of the tile's first line); K2 picks the bucket once
//  the phase is known.  Nothing per line is written at all.
//
// HBM traffic: input once + 8 B per line (written once, read once) + 28 B per
// tile of directory; the fused flavour: input once + 60 B per tile.
#include <cuda.h>

#include <atomic>

#include <type_traits>

#include "common.cuh"
#include "tma_tile.cuh"
#include "exon_b200_internal.h"
#include "x87div.h"
#include "combine_records.cuh"

namespace exb {

constexpr int FQ_WARPS = 4;     // warps per CTA (independent of each other)
constexpr int FQ_THREADS = FQ_WARPS * 32;
constexpr int REC_BLOCK = 512;  // records a warp reserves per bump allocation

// ---------------------------------------------------------------- tail word
// [63:62] 1 = no newline in the tile, 2 = has one
// [61:60] the line open after the last newline starts with '@' (2) / '+' (1) (valid if its start lies in the tile)
// [59:58] the tile's first byte is '@' / '+'
// [52:40] offset in the tile of the byte after its last newline (0..4096)
// [39:27] G/C count of the part after it          (0..4096)
// [20:0]  signed byte sum of that part            (|.| <= 2^19)
__device__ __forceinline__ uint64_t tail_pack(uint32_t state, uint32_t flags4, uint32_t rel, uint32_t g, int s) {
    return ((uint64_t)state << 62) | ((uint64_t)flags4 << 58) | ((uint64_t)rel << 40) | ((uint64_t)g << 27) |
           (uint64_t)((uint32_t)s & 0x1FFFFFu);
}
__device__ __forceinline__ int tail_s_of(uint64_t w) { return ((int)((uint32_t)w << 11)) >> 11; }
__device__ __forceinline__ uint32_t tail_g_of(uint64_t w) { return (uint32_t)(w >> 27) & 0x1FFFu; }
__device__ __forceinline__ uint32_t tail_rel_of(uint64_t w) { return (uint32_t)(w >> 40) & 0x1FFFu; }
__device__ __forceinline__ uint32_t tail_open_flags(uint64_t w) { return (uint32_t)(w >> 60) & 3u; }
__device__ __forceinline__ uint32_t tail_byte0_flags(uint64_t w) { return (uint32_t)(w >> 58) & 3u; }
__device__ __forceinline__ uint32_t at_plus_flags(int byte) { return (byte == '@' ? 2u : 0u) | (byte == '+' ? 1u : 0u); }

// round(gc_content * 2^32) of one sequence line, gc_content = (float)#GC / (float)len exactly as the scalar function
// computes it (sequence_functions/module.cpp:131-158); 0 for an empty line.  Same expression as seq_totals_kernel.
__device__ __forceinline__ unsigned long long gc_fix32(uint32_t g, uint32_t l) {
    return l ? (unsigned long long)__double2ll_rn((double)__fdiv_rn(__uint2float_rn(g), __uint2float_rn(l)) * 4294967296.0) : 0ull;
}

// ---------------------------------------------------------------- line record (8 bytes per newline)
// .x = byte-sum prefix at the newline (tile-relative, excludes the newline)
// .y = [11:0] position in the tile | [12] CR before it | [14:13] next line starts with '@' / '+' | [27:15] G/C prefix
__device__ __forceinline__ uint32_t rec_pack(int pos, uint32_t cr, uint32_t next_flags, int pg) {
    return (uint32_t)pos | (cr << 12) | (next_flags << 13) | ((uint32_t)pg << 15);
}
__device__ __forceinline__ int rec_pos(uint32_t y) { return (int)(y & 0xFFFu); }
__device__ __forceinline__ uint32_t rec_cr(uint32_t y) { return (y >> 12) & 1u; }
__device__ __forceinline__ uint32_t rec_next_flags(uint32_t y) { return (y >> 13) & 3u; }
__device__ __forceinline__ int rec_pg(uint32_t y) { return (int)(y >> 15); }

// ---------------------------------------------------------------- shared memory of one warp (K1)
// Tile data lives in two 4 KiB buffers per warp (1024-byte aligned: the 128B swizzle pattern is a function of
// address bits [7:9]); the small per-tile arrays follow all the data buffers.
constexpr int ENT_CAP = 64;  // newline entries a warp keeps per pass (a 4 KiB tile of 150 bp reads has ~45 newlines)
template <int FLAGS>
struct FqWarpAux {
    static constexpr bool kSeq = (FLAGS & EXB_F_SEQ) != 0, kQual = (FLAGS & EXB_F_QUAL) != 0;
    static constexpr int off_cpre = 0;                                   // int16[256]: ROW-relative byte-sum prefix at each 16-byte chunk (lane l owns [8l, 8l+8))
    static constexpr int off_ent = off_cpre + (kQual ? 256 * 2 : 0);     // uint2[ENT_CAP]: the tile's newlines in tile order {byte-sum prefix, packed word}
    static constexpr int off_bar = off_ent + ENT_CAP * 8;                // 2 mbarriers
    static constexpr int total = off_bar + 16;
    // the kernel has no static shared memory, so the dynamic region starts right after the 1 KiB the system
    // reserves per CTA: 1024-byte aligned, which the 128B swizzle needs (the kernel traps if that ever changes)
    static constexpr int off_wlut = FQ_WARPS * (2 * WT_BYTES + total);   // uint4[17], shared by the CTA
    static constexpr int off_preds = off_wlut + 17 * 16;                 // exb_predicate[EXB_MAX_PREDICATES] (fused flavour)
    static constexpr int off_nflut = off_preds + EXB_MAX_PREDICATES * (int)sizeof(exb_predicate);  // u8[256]: byte -> '@' (2) / '+' (1) flags
    static constexpr int cta_bytes = off_nflut + 256;
};

// single-predicate fast paths of the fused flavour, resolved once per kernel
struct FusedPlan {
    int simple;    // 1: exactly one EXB_P_MEAN_QUALITY predicate with op in {>, >=, <, <=}; 2: and c * 2^20 is a 32-bit integer
    double c;
    int want_pos;  // verdict = (d > 0) == want_pos when |d| is clear of rounding
    int op;
    int c20;       // simple == 2: c * 2^20
    int mul_q, mul_n;  // i32 != 0: e = sum * mul_q + n * mul_n has the sign of the verdict, exactly, in 32 bits
    int i32;
};

// the general form: any number of predicates, any constant (out of line, scalar arguments only: the hot loop
// inlines just the integer fast path).  `preds` points to shared memory.
__device__ __noinline__ bool fused_pass_general(const exb_predicate* preds, int n_preds, int simple, double c, int op, int want_pos, int qs,
                                                uint32_t len) {
    if (simple) {
        if (len == 0) return false;
        // d = sum - c n with ONE rounding.  |sum| < 2^20 inside a tile, so |d| > 1e-7 puts the exact quotient more
        // than 40 ulp from c whatever |c n| is (see exb_mean_cmp): neither rounding of the x87 path can cross.
        const double d = fma(-c, (double)len, (double)qs);
        if (fabs(d) > 1e-7) return (d > 0) == (want_pos != 0);
        return exb_mean_cmp_close((int64_t)qs, len, op, c);
    }
    bool ok = true;
    for (int i = 0; i < n_preds; i++) {
        const exb_predicate p = preds[i];
        ok = ok && (p.field == EXB_P_MEAN_QUALITY ? exb_mean_cmp((int64_t)qs, len, p.op, p.value) : exb_cmp((double)len, p.op, p.value));
    }
    return ok;
}
__device__ __forceinline__ bool fused_pass(const exb_predicate* preds, int n_preds, const FusedPlan& plan, int qs, uint32_t len) {
    if (plan.simple == 2) {
        // D = 2^20 (sum - c n), exact in 64-bit integers (|sum| < 2^20 and n <= 4096 inside a tile).  D != 0 means
        // |sum - c n| >= 2^-20 > 1e-7: the exact quotient is more than 40 ulp from c (see exb_mean_cmp), so neither
        // rounding of the x87 path can cross and the sign of D is the verdict.  (n = 0: D = 0 -> general form -> false.)
        const long long D = ((long long)qs << 20) - (long long)plan.c20 * (long long)len;
        if (D != 0) return (D > 0) == (plan.want_pos != 0);
    }
    return fused_pass_general(preds, n_preds, plan.simple, plan.c, plan.op, plan.want_pos, qs, len);
}
__device__ __forceinline__ FusedPlan make_plan(const FastqScanArgs& a) {
    FusedPlan p;
    p.simple = 0;
    p.c = 0;
    p.want_pos = 0;
    p.op = 0;
    p.c20 = 0;
    p.mul_q = p.mul_n = p.i32 = 0;
    if (a.n_fused == 1 && a.fused[0].field == EXB_P_MEAN_QUALITY && a.fused[0].op <= EXB_OP_LE) {
        p.simple = 1;
        p.c = a.fused[0].value;
        p.op = a.fused[0].op;
        p.want_pos = (p.op == EXB_OP_GT || p.op == EXB_OP_GE) ? 1 : 0;
        const double s = p.c * 1048576.0;
        if (fabs(s) < 2147483648.0 && s == rint(s)) {
            p.simple = 2;
            p.c20 = (int)s;
            // c = cq / 2^sh with sh <= 10 and |cq| < 2^18: 2^sh (sum - c n) = sum 2^sh - cq n stays below 2^31 for
            // |sum| < 2^20, n <= 4096 (one tile); multiplied by +-1 so that "e > 0" is "passes"
            int sh = 20, cq = p.c20;
            while (sh > 0 && (cq & 1) == 0) {
                cq >>= 1;
                sh--;
            }
            if (cq == 0) sh = 0;
            if (sh <= 10 && cq > -(1 << 18) && cq < (1 << 18)) {
                const int sgn = p.want_pos ? 1 : -1;
                p.i32 = 1;
                p.mul_q = sgn * (1 << sh);
                p.mul_n = -sgn * cq;
            }
        }
    }
    return p;
}

// shared-memory accessors on 32-bit shared-window addresses: the hot loop does no generic-pointer arithmetic
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// Read-only loads of data that does not change while a tile is being worked on (the tile buffer after its TMA fill, the
// LUTs): NOT volatile, so the scheduler may overlap the load -> IDP chains of independent newlines instead of keeping
// every shared-memory access in program order.  `tok` is a value produced by a volatile asm AFTER the buffer became
// readable (see the tile loop): the data dependence keeps the loads below that point, and a new value per tile keeps
// the compiler from reusing a result across tiles.
__device__ __forceinline__ uint4 lds128_ro(uint32_t addr, uint32_t tok) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "r"(tok));
    return v;
}
__device__ __forceinline__ uint32_t lds8_ro(uint32_t addr, uint32_t tok) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr), "r"(tok));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
// index of the most significant set bit (bfind) / its distance from bit 31 (= clz); 0xFFFFFFFF for 0 in both forms
__device__ __forceinline__ int bfind_u32(uint32_t x) {
    int r;
    asm("bfind.u32 %0, %1;\n" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t bfind_sh(uint32_t x) {
    uint32_t r;
    asm("bfind.shiftamt.u32 %0, %1;\n" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ int lds_s16(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};\n" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts64_if(bool on, uint32_t addr, uint32_t x, uint32_t y) {  // predicated, not branched around
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p st.shared.v2.u32 [%0], {%1,%2};\n}\n" ::"r"(addr), "r"(x), "r"(y), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t x) {
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(addr), "r"(x) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, int x, int y, int z, int w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// inclusive warp scan; the shuffle's predicate output (source lane in range) guards the add: two instructions per step
__device__ __forceinline__ uint32_t warp_incl_scan_u32p(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n .reg .pred p;\n .reg .u32 t;\n shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n @p add.u32 %0, %0, t;\n}\n" : "+r"(v) : "r"(d));
    return v;
}
// tile-local byte index -> byte offset in the 128B-swizzled tile buffer (same map as sidx, two instructions)
__device__ __forceinline__ uint32_t swz(uint32_t li) { return li ^ ((li >> 3) & 0x70u); }

// =================================================================== K1
template <int FLAGS>
__global__ void __launch_bounds__(FQ_THREADS, 6) fastq_tile_kernel(const __grid_constant__ CUtensorMap tmap, const FastqScanArgs a,
                                                                const uint32_t c7f, const uint32_t c7b) {
    using AUX = FqWarpAux<FLAGS>;
    constexpr bool kSeq = AUX::kSeq, kQual = AUX::kQual;
    constexpr bool kFused = (FLAGS & EXB_F_FUSED) != 0;
    // fused TOTALS flavour (C5: COUNT, SUM(len), SUM(#GC), AVG(gc_content) of the sequence lines): no predicate, the
    // aggregates of every line that is a SEQUENCE line under one of the four hypotheses
    constexpr bool kFusedSeq = kFused && kSeq && !kQual;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint4* s_wlut = reinterpret_cast<uint4*>(smem_raw + AUX::off_wlut);  // s_wlut[k]: 0x01 in the first k bytes -- IDP.4A weights

    // the warp index through a shuffle: the compiler then knows it (and every tile index / buffer address derived from it)
    // is warp-uniform and keeps that arithmetic -- and the TMA descriptor operands -- in uniform registers
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x < 17) {
        const int k = threadIdx.x;
        auto w = [k](int b) -> uint32_t {  // word holding bytes [b, b+4)
            const int c = k - b;
            return c <= 0 ? 0u : (c >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - c))));
        };
        s_wlut[k] = make_uint4(w(0), w(4), w(8), w(12));
    }
    exb_predicate* s_preds = reinterpret_cast<exb_predicate*>(smem_raw + AUX::off_preds);
    if (kFused && threadIdx.x < a.n_fused) s_preds[threadIdx.x] = a.fused[threadIdx.x];
    for (int i = threadIdx.x; i < 256; i += FQ_THREADS) (smem_raw + AUX::off_nflut)[i] = (uint8_t)at_plus_flags(i);

    uint32_t smem_u32;  // volatile: computed once and kept in a register (the compiler otherwise re-derives it ~5 times per tile)
    asm volatile("{\n .reg .u64 t;\n cvta.to.shared.u64 t, %1;\n cvt.u32.u64 %0, t;\n}\n" : "=r"(smem_u32) : "l"(smem_raw));
    if ((smem_u32 & 1023u) != 0) __trap();
    const uint32_t data0_u32 = smem_u32 + warp * (2 * WT_BYTES);
    const uint32_t aux_u32 = smem_u32 + FQ_WARPS * (2 * WT_BYTES) + warp * AUX::total;
    const uint32_t cpre_u32 = aux_u32 + AUX::off_cpre + lane * 16;  // this lane's eight chunk prefixes (int16, relative to its row)
    const uint32_t ent_u32 = aux_u32 + AUX::off_ent;
    const uint32_t bar0 = aux_u32 + AUX::off_bar;
    const uint32_t wlut_u32 = smem_u32 + AUX::off_wlut, nflut_u32 = smem_u32 + AUX::off_nflut;
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();  // the only block-wide barrier of the kernel (LUTs + barriers ready)

    const uint8_t* __restrict__ buf = a.buf;
    const int64_t origin = a.begin & ~(int64_t)15;
    // tile indices are 32-bit in the loop (the launcher refuses inputs of 2^31 tiles = 8 TiB)
    const int n_tiles = (int)a.n_tiles;
    const int64_t full_rows = a.tma_rows;  // rows of 128 bytes that lie completely inside [origin, n): what the tensor map covers
    const int first_edge = (int)(full_rows / WT_ROWS);  // first tile with rows the tensor map does not cover
    const int tma_tiles = (int)((full_rows + WT_ROWS - 1) / WT_ROWS) < n_tiles ? (int)((full_rows + WT_ROWS - 1) / WT_ROWS) : n_tiles;
    const bool pad0 = a.begin != origin;
    const uint32_t pat_nl = c7f & 0x0A0A0A0Au, pat_gc = c7f & 0x43434343u;  // derived from an argument: stay in registers
    const int stride = (int)gridDim.x * FQ_WARPS;
    const FusedPlan plan = kFused ? make_plan(a) : FusedPlan{0, 0.0, 0, 0, 0, 0, 0, 0};
    const int lane_off = lane * ROW_BYTES;
    const uint32_t lane_part = (uint32_t)lane_off | ((uint32_t)(lane & 7) << 4);  // chunk c of the lane's row sits at (buffer + lane_part) ^ (c << 4)
    // the bytes just outside the lane's row: last byte of row lane - 1 / first byte of row lane + 1 (offsets in the swizzled
    // tile buffer), clamped into the tile
    const uint32_t o_prev = lane == 0 ? lane_part : swz((uint32_t)(lane_off - 1));
    const uint32_t o_next = lane == 31 ? (lane_part ^ 127u) : swz((uint32_t)(lane_off + ROW_BYTES));
    // fused: what the first byte of the line after newline i says about the hypotheses (i = lane mod 4, see B2)
    const uint32_t bad_hdr = 1u << ((0 - (lane + 1)) & 3), bad_plus = 1u << ((2 - (lane + 1)) & 3);
    const uint32_t bad_lut = (bad_hdr | bad_plus) | (bad_hdr << 4) | (bad_plus << 8);

    // ---- staging of one tile into buffer b (asynchronous; one instruction from one lane)
    auto issue = [&](int tile, int b) {
        if (tile < tma_tiles && lane == 0) {
            mbar_expect_tx(bar0 + 8 * b, WT_BYTES);
            tma_load_tile(data0_u32 + b * WT_BYTES, &tmap, tile * WT_ROWS, bar0 + 8 * b);
        }
    };

    // bump allocation of record space, REC_BLOCK at a time per warp
    int64_t rec_next = 0;
    int rec_left = 0;  // a warp that has not allocated yet owns nothing: a record is never written through off0 then
    bool blk_ok = true;  // the current block lies inside the record space

    int cur = (int)blockIdx.x * FQ_WARPS + warp;
    issue(cur, 0);
    int b = 0;
    uint32_t phase_bits = 0;  // bit b = parity the next wait on buffer b expects

    while (cur < n_tiles) {
        const int nxt = cur + stride;
        issue(nxt, b ^ 1);

        const int tile = cur;
        const uint32_t sb = data0_u32 + b * WT_BYTES;
        if (tile < tma_tiles) {
            const uint32_t par = (phase_bits >> b) & 1u;
            while (!mbar_try_wait(bar0 + 8 * b, par)) {
            }
            phase_bits ^= 1u << b;
        }
        // Rare edge tiles (uniform per warp): rows the tensor map does not cover, bytes before `begin`, the virtual '\n'
        uint32_t b0_addr = sb;           // where the tile's first parsed byte sits (not byte 0 if the range starts inside tile 0's first 16 bytes)
        int virt = -1;                   // tile-local position of the virtual '\n' that terminates an unterminated last line
        int last_known = WT_BYTES - 1;   // first tile-local position whose NEXT byte lies outside the tile or the parse range
        if (tile >= first_edge || (tile == 0 && pad0)) {
            const int64_t tile_base = origin + (int64_t)tile * WT_BYTES;
            const int64_t row0 = (int64_t)tile * WT_ROWS;
            uint8_t* sbytes = smem_raw + (sb - smem_u32);
            const bool partial = tile >= first_edge;
            if (row0 >= full_rows) {  // nothing came through TMA: clear the buffer
                for (int i = lane; i < WT_BYTES / 16; i += 32) reinterpret_cast<uint4*>(sbytes)[i] = make_uint4(0, 0, 0, 0);
                __syncwarp();
            }
            if (partial) {  // bytes of the last, incomplete row (TMA zero-filled everything beyond the tensor)
                const int64_t tail0 = origin + full_rows * ROW_BYTES;
                const int64_t lo = tail0 > tile_base ? tail0 : tile_base;
                for (int64_t g = lo + lane; g < a.n && g < tile_base + WT_BYTES; g += 32) sbytes[sidx((int)(g - tile_base))] = buf[g];
            }
            __syncwarp();
            if (lane == 0) {
                if (tile == 0 && pad0)
                    for (int64_t i = origin; i < a.begin; i++) sbytes[sidx((int)(i - origin))] = 0;
                if (a.is_final && a.n >= tile_base && a.n < tile_base + WT_BYTES) {  // an unterminated last line gets a virtual '\n' at n
                    const bool open = a.n > a.begin ? buf[a.n - 1] != '\n' : (a.prev && a.prev->open_line_start < a.n);
                    if (open) sbytes[sidx((int)(a.n - tile_base))] = '\n';
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes before this buffer's next TMA fill
            __syncwarp();
            if (tile == 0 && pad0) b0_addr = sb + swz((uint32_t)(a.begin - origin));
            if (a.is_final && a.n >= tile_base && a.n < tile_base + WT_BYTES) virt = (int)(a.n - tile_base);
            if (partial && a.n - tile_base < WT_BYTES) last_known = (int)(a.n - tile_base) - 1;
        }

        // the buffer is readable from here on: the read-only loads below hang on this token (see lds128_ro)
        uint32_t tok;
        asm volatile("mov.u32 %0, %1;\n" : "=r"(tok) : "r"(tile) : "memory");

        // ---- A. analysis of the lane's row: two 64-byte halves
        uint32_t w_nl[4];  // newline mask of the row, 32 bytes per word
        uint64_t gm[2] = {0, 0};
        int ex_cnt, n_events, cnt, ex_s = 0, total_s = 0, ex_g = 0, total_g = 0, g0 = 0;
        const uint32_t rowx = sb + lane_part;
        {
            int acc = 0;
            int pre[8];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint4 c0 = lds128_ro(rowx ^ ((4 * h + 0) << 4), tok), c1 = lds128_ro(rowx ^ ((4 * h + 1) << 4), tok),
                            c2 = lds128_ro(rowx ^ ((4 * h + 2) << 4), tok), c3 = lds128_ro(rowx ^ ((4 * h + 3) << 4), tok);
                w_nl[2 * h] = nl_mask32r(c0, c1, c7f, pat_nl);
                w_nl[2 * h + 1] = nl_mask32r(c2, c3, c7f, pat_nl);
                if (kQual) {
                    pre[4 * h + 0] = acc;
                    acc = sbyte_sum16(c0, acc);
                    pre[4 * h + 1] = acc;
                    acc = sbyte_sum16(c1, acc);
                    pre[4 * h + 2] = acc;
                    acc = sbyte_sum16(c2, acc);
                    pre[4 * h + 3] = acc;
                    acc = sbyte_sum16(c3, acc);
                }
                if (kSeq)
                    gm[h] = ((uint64_t)gc_mask32r(c2, c3, c7b, c7f, pat_gc) << 32) | gc_mask32r(c0, c1, c7b, c7f, pat_gc);
            }
            cnt = __popc(w_nl[0]) + __popc(w_nl[1]) + __popc(w_nl[2]) + __popc(w_nl[3]);
            g0 = kSeq ? __popcll(gm[0]) : 0;
            const int g1 = kSeq ? __popcll(gm[1]) : 0;
            if (kQual && !kSeq) {
                // ONE scan: newline count << 20 | byte sum, the sum biased by 2^14 per lane so that it is never negative
                // (a row of 128 signed bytes sums to [-2^14, 2^14); 32 rows stay below 2^20)
                const uint32_t packed = ((uint32_t)cnt << 20) + (uint32_t)(acc + 16384);
                const uint32_t incl = warp_incl_scan_u32p(packed);
                const uint32_t ex = incl - packed;
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                ex_cnt = (int)(ex >> 20);
                n_events = (int)(tot >> 20);
                ex_s = (int)(ex & 0xFFFFFu) - 16384 * lane;
                total_s = (int)(tot & 0xFFFFFu) - 16384 * 32;
            } else {
                // newlines and G/C (each <= 4096 per tile: 13 bits) share one scan word; the byte sums get their own
                const uint32_t packed = ((uint32_t)cnt << 16) + (uint32_t)(g0 + g1);
                const uint32_t incl = warp_incl_scan_u32p(packed);
                const uint32_t ex = incl - packed;
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                ex_cnt = (int)(ex >> 16);
                n_events = (int)(tot >> 16);
                ex_g = (int)(ex & 0xFFFFu);
                total_g = (int)(tot & 0xFFFFu);
                if (kQual) {
                    const uint32_t si = warp_incl_scan_u32p((uint32_t)acc);
                    ex_s = (int)(si - (uint32_t)acc);
                    total_s = (int)__shfl_sync(0xffffffffu, si, 31);
                }
            }
            if (kQual) {
                // read back by this lane only (dynamic index at each newline): no warp barrier needed.  Row-relative, so
                // 16 bits are enough (|sum of 128 signed bytes| <= 2^14); pre[0] = 0.
                sts128(cpre_u32, (int)__byte_perm((uint32_t)pre[0], (uint32_t)pre[1], 0x5410), (int)__byte_perm((uint32_t)pre[2], (uint32_t)pre[3], 0x5410),
                       (int)__byte_perm((uint32_t)pre[4], (uint32_t)pre[5], 0x5410), (int)__byte_perm((uint32_t)pre[6], (uint32_t)pre[7], 0x5410));
            }
        }
        if (lane == 0) a.tile_cnt[tile] = (uint32_t)n_events;

        // record space for this tile's newlines: the warp's current block first, the rest in a fresh block
        // (a tile's records are at most two runs, so no slot of a block is ever abandoned)
        int64_t rec_off0 = 0, rec_off1 = 0;
        int rec_n0 = 0;
        const bool rec_ok0 = blk_ok;  // the first run lives in the block allocated earlier
        if (!kFused) {
            rec_off0 = rec_next;
            rec_n0 = n_events < rec_left ? n_events : rec_left;
            rec_next += rec_n0;
            rec_left -= rec_n0;
            const int rest = n_events - rec_n0;
            if (rest > 0) {
                const int need = rest > REC_BLOCK ? rest : REC_BLOCK;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.rec_bump, (unsigned long long)need);
                rec_off1 = (int64_t)__shfl_sync(0xffffffffu, base, 0);
                rec_next = rec_off1 + rest;
                rec_left = need - rest;
                blk_ok = rec_off1 + need <= a.rec_space;
                if (!blk_ok && lane == 0) a.result->overflow = 1;
            }
            if (lane == 0) {
                a.tile_rec[tile] = rec_off0;
                a.tile_rec2[tile] = rec_off1 | ((int64_t)rec_n0 << 48);
            }
        }

        // ---- B. the tile's newlines, in two steps.
        // B1 (row frame): every lane files the newlines of ITS OWN row -- {byte-sum prefix, packed word} of the k-th
        //    one under its tile-local line index ex_cnt + k -- in the warp's entry table.  Branch-free: a lane without
        //    a k-th newline computes on a harmless in-range position and stores nothing.  Rows of real FASTQ hold at
        //    most two newlines ("...\n+\n..."), so the common case needs no bit-clearing walk over the 128-bit row mask:
        //    the first newline is a min and the last one a max over four per-word find-first / find-last results.
        // B2 (tile frame): lane l takes entries l, l + 32, ...: line i is what lies between newline i - 1 and newline i,
        //    whichever rows those are in.  Fused: line i is a quality line under exactly one phase hypothesis of the
        //    tile's line 0, h = (3 - i) & 3 = (3 - lane) & 3 -- ONE bucket per lane, fixed for the whole kernel.
        //    General: entry i is record i of the tile (coalesced 8-byte stores).
        // The table holds ENT_CAP entries; a tile with more newlines (reads shorter than ~60 bp) repeats B1 + B2 per
        // ENT_CAP indices.  The tile's line 0 began in an earlier tile: K2 finishes it.
        uint32_t f_cq = 0;   // lines of this lane's bucket that pass: count | length sum << 12
        int f_qs = 0;        // their Phred sums
        uint32_t f_bad = 0;  // bit h: a line start contradicts hypothesis h
        unsigned long long f_fix = 0;  // totals flavour: sum of round(gc_content * 2^32) over the bucket's lines
        uint2 e_first = make_uint2(0u, 0u), e_last = make_uint2(0u, 0u), carry = make_uint2(0u, 0u);
        const int rounds = __reduce_max_sync(0xffffffffu, cnt);
        const uint32_t a_prev = sb + o_prev, a_next = sb + o_next;

        // B1 for one newline: row position p (0..127), the lane's k-th, table slot ex_cnt + k - base
        auto file_entry = [&](int k, int p, int base, auto one_pass) {
            const int pos = lane_off + p;
            int ps = 0, pg = 0;
            if (kQual) {
                const uint4 v = lds128_ro(rowx ^ (uint32_t)(p & 0x70), tok);
                const uint4 w = lds128_ro(wlut_u32 + (uint32_t)((p & 15) << 4), tok);
                const int cp = lds_s16(cpre_u32 | (uint32_t)((p >> 3) & 0xE));  // written by this lane above: stays in program order
                int acc = __dp4a((int)v.x, (int)w.x, ex_s);
                acc = __dp4a((int)v.y, (int)w.y, acc);
                acc = __dp4a((int)v.z, (int)w.z, acc);
                acc = __dp4a((int)v.w, (int)w.w, acc);
                ps = acc + cp;
            }
            if (kSeq) pg = p < 64 ? ex_g + __popcll(gm[0] & low_bits64(p)) : ex_g + g0 + __popcll(gm[1] & low_bits64(p - 64));
            // neighbours: byte q of the lane's own row sits at rowx ^ q (rowx is 16-byte aligned and the swizzle is an XOR
            // of bits 4..6); the byte before the row / after it at a per-lane offset of the buffer (lane 0 / 31: clamped
            // into the row; the byte before the tile's first byte is patched in below, the byte after its last one gets
            // no verdict: last_known)
            const uint32_t before = lds8_ro(p == 0 ? a_prev : (rowx ^ (uint32_t)(p - 1)), tok);
            const uint32_t after = lds8_ro(p == 127 ? a_next : (rowx ^ (uint32_t)(p + 1)), tok);
            // a CR directly before a real LF is stripped; the virtual '\n' at EOF strips nothing
            const uint32_t cr = (before == '\r' && pos != virt) ? 1u : 0u;
            // '@' / '+' flags of the next line's first byte; 3 = that byte is not in this tile or range: no verdict here
            const uint32_t nfk = lds8_ro(nflut_u32 + after, tok);
            const uint32_t nf = pos < last_known ? nfk : 3u;
            const uint32_t idx = (uint32_t)(ex_cnt + k - base);
            sts64_if(k < cnt && (decltype(one_pass)::value || idx < (uint32_t)ENT_CAP), ent_u32 + idx * 8u, (uint32_t)ps, rec_pack(pos, cr, nf, kSeq ? pg : 0));
        };
        // the byte before the tile's first byte lives in global memory: a newline at position 0 of the tile (entry 0,
        // filed by lane 0) learns its CR flag here (rare, so it is kept out of file_entry)
        auto patch_entry0 = [&]() {
            if (lane == 0 && (w_nl[0] & 1u) && virt != 0) {
                const int64_t tile_base = origin + (int64_t)tile * WT_BYTES;
                if (tile_base - 1 >= (a.prev ? 0 : a.begin) && buf[tile_base - 1] == '\r') sts32(ent_u32 + 4u, lds32(ent_u32 + 4u) | (1u << 12));
            }
        };
        // B2 for entries i0 + lane of a pass that starts at tile-local index `base` and holds nb entries.  Returns whether
        // this lane's line sits exactly ON the threshold (rare: the two roundings of the reference decide, see `settle`).
        auto lines = [&](int i0, int nb, int base, auto one_pass, int& qs, uint32_t& len) -> bool {
            const int i = i0 + lane;
            const bool valid = i < nb;
            const uint2 e = lds64(ent_u32 + (uint32_t)i * 8u);  // i < ENT_CAP: in range even when not valid
            uint2 pe = lds64(ent_u32 + (uint32_t)(i > 0 ? i - 1 : 0) * 8u);
            if (!decltype(one_pass)::value && i == 0) pe = carry;  // the last newline of the previous pass (unused for the tile's line 0)
            if (!kFused) {
                const int t = base + i;
                if (valid && (t < rec_n0 ? rec_ok0 : blk_ok)) a.records[t < rec_n0 ? rec_off0 + t : rec_off1 + (t - rec_n0)] = e;
                return false;
            } else {
                // the line that STARTS after newline i is line i + 1 of the tile: a header under h = -(i + 1), a plus
                // line under h = 2 - (i + 1).  nf: 0 -> both contradicted, 1 ('+') -> header, 2 ('@') -> plus, 3 -> none
                f_bad |= (bad_lut >> (4u * (valid ? rec_next_flags(e.y) : 3u))) & 15u;
                // line i: between newline i - 1 and newline i.  A CR flag implies a non-empty line (the byte before an
                // empty line's newline is the previous newline), so the length needs no clamp.
                const bool line = valid && (i | base) != 0;
                const uint32_t cr = rec_cr(e.y);
                len = (uint32_t)(rec_pos(e.y) - rec_pos(pe.y) - 1) - cr;
                if (kFusedSeq) {
                    // line i is a sequence line under h = (1 - i) & 3 = (1 - lane) & 3: this lane's ONE bucket
                    const int gsum = rec_pg(e.y) - rec_pg(pe.y);
                    if (line) {
                        f_cq += 1u + (len << 12);
                        f_qs += gsum;
                        f_fix += gc_fix32((uint32_t)gsum, len);
                    }
                    return false;
                }
                qs = (int)e.x - (int)pe.x - 10 - 13 * (int)cr - 33 * (int)len;
                if (plan.i32) {
                    // ee = +-2^sh (sum - c n), exact in 32 bits; ee != 0 means |sum - c n| >= 2^-20 > 1e-7, which puts the
                    // exact quotient more than 40 ulp from c (see exb_mean_cmp): the sign of ee is the verdict.  n = 0: ee = 0.
                    const int ee = qs * plan.mul_q + (int)len * plan.mul_n;
                    if (line && ee > 0) {
                        f_cq += 1u + (len << 12);
                        f_qs += qs;
                    }
                    return line && ee == 0;
                }
                if (line && fused_pass(s_preds, a.n_fused, plan, qs, len)) {
                    f_cq += 1u + (len << 12);
                    f_qs += qs;
                }
                return false;
            }
        };
        auto settle = [&](bool close, int qs, uint32_t len) {
            if (close && fused_pass_general(s_preds, a.n_fused, plan.simple, plan.c, plan.op, plan.want_pos, qs, len)) {
                f_cq += 1u + (len << 12);
                f_qs += qs;
            }
        };

        if (rounds <= 2) {  // warp-uniform; at most 64 newlines: one pass
            // first newline of the row: ctz per word (0xFFFFFFFF for an empty word survives the OR), unsigned min;
            // last newline of the row: find-last per word (-1 for an empty word survives the OR), signed max.
            // Two rounds (the common case) run as ONE straight-line block so that their load -> IDP chains overlap.
            const uint32_t c0 = bfind_sh(__brev(w_nl[0])), c1 = bfind_sh(__brev(w_nl[1])) | 32u, c2 = bfind_sh(__brev(w_nl[2])) | 64u,
                           c3 = bfind_sh(__brev(w_nl[3])) | 96u;
            const int p_first = (int)(min(min(c0, c1), min(c2, c3)) & 127u);
            if (rounds == 2) {
                const int d0 = bfind_u32(w_nl[0]), d1 = bfind_u32(w_nl[1]) | 32, d2 = bfind_u32(w_nl[2]) | 64, d3 = bfind_u32(w_nl[3]) | 96;
                const int p_last = max(max(d0, d1), max(d2, d3)) & 127;
                file_entry(0, p_first, 0, std::true_type());
                file_entry(1, p_last, 0, std::true_type());
            } else if (rounds == 1) {
                file_entry(0, p_first, 0, std::true_type());
            }
            patch_entry0();
            __syncwarp();
            int qs0 = 0, qs1 = 0;
            uint32_t len0 = 0, len1 = 0;
            if (n_events > 32) {
                const bool close0 = lines(0, n_events, 0, std::true_type(), qs0, len0);
                const bool close1 = lines(32, n_events, 0, std::true_type(), qs1, len1);
                if (kFused && __any_sync(0xffffffffu, close0 || close1)) {
                    settle(close0, qs0, len0);
                    settle(close1, qs1, len1);
                }
            } else if (n_events > 0) {
                const bool close0 = lines(0, n_events, 0, std::true_type(), qs0, len0);
                if (kFused && __any_sync(0xffffffffu, close0)) settle(close0, qs0, len0);
            }
            e_first = lds64(ent_u32);
            e_last = lds64(ent_u32 + (uint32_t)(n_events > 0 ? n_events - 1 : 0) * 8u);
        } else {
            for (int base = 0;; base += ENT_CAP) {
                uint64_t mlo = ((uint64_t)w_nl[1] << 32) | w_nl[0], mhi = ((uint64_t)w_nl[3] << 32) | w_nl[2];
                for (int k = 0; k < rounds; k++) {
                    const bool use_hi = mlo == 0;  // also when no newline is left: p = 63 then
                    const uint64_t m = use_hi ? mhi : mlo;
                    const int p = __ffsll((long long)m) - 1 + (use_hi ? 64 : 0);  // position in the row
                    mlo &= mlo - 1;
                    mhi = use_hi ? (mhi & (mhi - 1)) : mhi;
                    file_entry(k, p, base, std::false_type());
                }
                if (base == 0) patch_entry0();
                __syncwarp();
                const int nb = n_events - base < ENT_CAP ? n_events - base : ENT_CAP;
                for (int i0 = 0; i0 < nb; i0 += 32) {
                    int qs0 = 0;
                    uint32_t len0 = 0;
                    const bool close0 = lines(i0, nb, base, std::false_type(), qs0, len0);
                    if (kFused && __any_sync(0xffffffffu, close0)) settle(close0, qs0, len0);
                }
                if (base == 0) e_first = lds64(ent_u32);  // the tile's first newline: K2 finishes the line it ends
                if (base + ENT_CAP >= n_events) {
                    e_last = lds64(ent_u32 + (uint32_t)(nb - 1) * 8u);  // the tile's last newline
                    break;
                }
                carry = lds64(ent_u32 + (uint32_t)(ENT_CAP - 1) * 8u);
                __syncwarp();  // everybody has read the table before the next pass overwrites it
            }
        }

        // ---- C. tail word: what follows the tile's last newline (local information only)
        {
            const uint32_t b0 = lds8_ro(nflut_u32 + lds8_ro(b0_addr, tok), tok);
            uint64_t tw;
            if (n_events > 0) {
                tw = tail_pack(2, (rec_next_flags(e_last.y) << 2) | b0, (uint32_t)(rec_pos(e_last.y) + 1), (uint32_t)(total_g - rec_pg(e_last.y)),
                               total_s - ((int)e_last.x + 10));
            } else {
                tw = tail_pack(1, b0, 0, (uint32_t)total_g, total_s);
            }
            if (lane == 0) a.tails[tile] = tw;
        }

        if (kFused) {
            // lanes l, l + 4, ... share bucket (3 - l) & 3: three butterfly steps leave its totals in lanes 0..3
#pragma unroll
            for (int d = 4; d < 32; d <<= 1) {
                f_cq += __shfl_xor_sync(0xffffffffu, f_cq, d);
                f_qs += __shfl_xor_sync(0xffffffffu, f_qs, d);
                if (kFusedSeq) f_fix += __shfl_xor_sync(0xffffffffu, f_fix, d);
            }
            const uint32_t bad4 = __reduce_or_sync(0xffffffffu, f_bad);
            FusedTile* ft = a.fused_tiles + tile;
            if (lane < 4) {
                const int bk = kFusedSeq ? ((1 - lane) & 3) : ((3 - lane) & 3);  // sequence lines: h = 1 - i; quality lines: h = 3 - i
                ft->cq[bk] = f_cq;
                ft->qs[bk] = f_qs;
                if (kFusedSeq) a.fused_fix[(int64_t)tile * 4 + bk] = f_fix;
            }
            if (lane == 4) *reinterpret_cast<uint4*>(&ft->ps0) = make_uint4(e_first.x, e_first.y, bad4, 0u);
        }

        __syncwarp();  // every lane is done with this buffer before lane 0 lets TMA refill it
        cur = nxt;
        b ^= 1;
    }
}

// =================================================================== K2 helpers
struct OpenLine {  // the line that is open at a tile's first byte
    int64_t start;
    int s, g;        // byte sum / G,C count of its part before the tile (mod 2^32)
    uint32_t flags;  // first byte is '@' (bit 1) / '+' (bit 0)
};

// Walk the predecessors' tail words back to the tile that holds the open line's start: tile-1 unless
// lines are longer than a tile.  Serial per calling thread (K1 has finished: plain loads, no polling).
__device__ __forceinline__ OpenLine open_line_before(const uint64_t* __restrict__ tails, int64_t tile, int64_t origin, const FastqScanArgs& a) {
    OpenLine o;
    o.s = 0;
    o.g = 0;
    uint32_t next_b0 = tail_byte0_flags(tails[tile]);  // first-byte flags of the tile after the one being inspected
    for (int64_t k = tile - 1; k >= 0; k--) {
        const uint64_t w = tails[k];
        o.s += tail_s_of(w);
        o.g += (int)tail_g_of(w);
        if ((uint32_t)(w >> 62) == 2u) {
            const uint32_t rel = tail_rel_of(w);
            o.start = origin + k * WT_BYTES + rel;
            o.flags = rel < (uint32_t)WT_BYTES ? tail_open_flags(w) : next_b0;  // rel = 4096: the line starts with the next tile
            return o;
        }
        next_b0 = tail_byte0_flags(w);
    }
    if (a.prev && a.prev->open_line_start < a.begin) {  // the line began in an earlier range of a chained scan
        o.start = a.prev->open_line_start;
        o.s += (int)a.prev->tail_s;
        o.g += (int)a.prev->tail_g;
        o.flags = a.prev->pad & 3u;
    } else {  // it starts at `begin`: tile 0's byte-0 flags describe that byte
        o.start = a.begin;
        o.flags = next_b0;
    }
    return o;
}

__device__ __forceinline__ void write_final_state(const FastqScanArgs& a, uint64_t total_lines, int n_events, int64_t tile_base, uint64_t tw,
                                                  const OpenLine& open) {
    a.result->total_lines = total_lines;
    if (n_events > 0) {
        a.result->open_line_start = tile_base + tail_rel_of(tw);
        a.result->tail_s = tail_s_of(tw);
        a.result->tail_g = tail_g_of(tw);
        a.result->pad = tail_open_flags(tw);  // meaningless if the open line has no byte yet (it then starts at the next range's begin)
    } else {
        a.result->open_line_start = open.start;
        a.result->tail_s = (int64_t)open.s + tail_s_of(tw);
        a.result->tail_g = (int64_t)open.g + tail_g_of(tw);
        a.result->pad = open.flags;
    }
}

// =================================================================== K2
// One warp per tile: turns the tile's records into per-line / per-record outputs.
template <typename OffT, int FLAGS>
__global__ void __launch_bounds__(256) fastq_emit_kernel(const FastqScanArgs a) {
    constexpr bool kLines = (FLAGS & EXB_F_LINES) != 0;
    constexpr bool kSeq = (FLAGS & EXB_F_SEQ) != 0, kQual = (FLAGS & EXB_F_QUAL) != 0;
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = a.n_tiles;
    const int64_t origin = a.begin & ~(int64_t)15;
    const uint64_t init = a.prev ? a.prev->total_lines : 0ull;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool overflowed = a.result->overflow != 0;  // K1 ran out of record space: the caller retries with more
    const uint64_t rec0 = a.local_records ? (init >> 2) : 0ull;  // byte-range shards index their records from the first one they touch

    // A warp takes 32 consecutive tiles at a time: lane l fetches tile (batch + l)'s directory words in ONE round of
    // coalesced loads, then the tiles are processed one after the other with the words broadcast by shuffles.  (One tile
    // per warp iteration paid three dependent global latencies per tile -- count, directory, records -- ~100 us of the
    // kernel's 164 us for a 1.4 GB file.)
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t n_crlf = 0;  // lines of this lane that ended in CR LF (the field split skips its CR probes when the range has none)
    for (int64_t batch = warp_global * 32; batch < n_tiles; batch += warps * 32) {
      const int64_t mine = batch + lane;
      const bool mv = mine < n_tiles;
      const int m_cnt = mv ? (int)a.tile_cnt[mine] : 0;
      const unsigned long long m_base = mv ? (unsigned long long)a.line_base[mine] : 0ull;
      const long long m_rec = mv ? a.tile_rec[mine] : 0ll, m_rec2 = mv ? a.tile_rec2[mine] : 0ll;
      const unsigned long long m_tail = mv ? a.tails[mine] : 0ull, m_tail_prev = (mv && mine > 0) ? a.tails[mine - 1] : 0ull;
      if (batch == 0 && lane == 0 && a.prev) {  // chained range: errors of earlier ranges stay visible in the last result
          if (a.prev->err_pos != 0ull) atomicMax(&a.result->err_pos, a.prev->err_pos);
          if (a.prev->overflow) a.result->overflow = 1;
      }
      uint32_t todo = __ballot_sync(0xffffffffu, mv && (m_cnt > 0 || mine == n_tiles - 1));
      while (todo) {
        const int j = __ffs((int)todo) - 1;
        todo &= todo - 1;
        const int64_t tile = batch + j;
        const int n_events = __shfl_sync(0xffffffffu, m_cnt, j);
        const bool is_last = tile == n_tiles - 1;
        const int64_t tile_base = origin + tile * WT_BYTES;
        const uint64_t excl = init + (uint64_t)__shfl_sync(0xffffffffu, m_base, j);
        const uint64_t tw = __shfl_sync(0xffffffffu, m_tail, j), twp = __shfl_sync(0xffffffffu, m_tail_prev, j);
        OpenLine open;
        if (tile > 0 && (uint32_t)(twp >> 62) == 2u) {  // the previous tile has a newline: the open line starts right there
            const uint32_t rel = tail_rel_of(twp);
            open.s = tail_s_of(twp);
            open.g = (int)tail_g_of(twp);
            open.start = origin + (tile - 1) * WT_BYTES + rel;
            open.flags = rel < (uint32_t)WT_BYTES ? tail_open_flags(twp) : tail_byte0_flags(tw);
        } else {
            open = open_line_before(a.tails, tile, origin, a);  // same addresses in every lane: broadcast loads
        }
        if (is_last && lane == 0) write_final_state(a, excl + (uint64_t)n_events, n_events, tile_base, tw, open);
        if (overflowed) continue;

        const int64_t off0 = __shfl_sync(0xffffffffu, m_rec, j), w1 = __shfl_sync(0xffffffffu, m_rec2, j);
        const int64_t off1 = w1 & 0xFFFFFFFFFFFFll;
        const int n0 = (int)(w1 >> 48);
        int c_ps = 0;
        uint32_t c_y = 0;
        // the per-line work of one record word: r = this line's newline, (pps, py) = the previous newline of the tile
        auto emit_line = [&](bool active, int li, uint2 r, int pps, uint32_t py) {
            if (active) {
                const uint64_t g = excl + (uint64_t)li;  // global index of the line this newline ends
                const int pos = rec_pos(r.y);
                const int64_t e = tile_base + pos;
                uint32_t len, first_flags;
                int ssum = 0, gsum = 0;
                int64_t line_start;
                if (li > 0) {
                    const int ppos = rec_pos(py);
                    len = (uint32_t)(pos - ppos - 1);
                    if (kQual) ssum = (int)r.x - pps - 10;
                    if (kSeq) gsum = rec_pg(r.y) - rec_pg(py);
                    line_start = tile_base + ppos + 1;
                    first_flags = rec_next_flags(py);
                } else {  // the line that was open when the tile began
                    line_start = open.start;
                    len = (uint32_t)(tile_base - open.start) + (uint32_t)pos;
                    if (kQual) ssum = (int)r.x + open.s;
                    if (kSeq) gsum = rec_pg(r.y) + open.g;
                    first_flags = open.flags;
                }
                if (g < a.max_lines) {
                    const uint32_t cr = len > 0 ? rec_cr(r.y) : 0u;
                    len -= cr;
                    n_crlf += cr;
                    const int ph = (int)(g & 3);
                    const uint64_t rix = (g >> 2) - rec0;
                    if (kLines) {
                        if (g < (uint64_t)a.line_cap) reinterpret_cast<OffT*>(a.line_end)[g] = (OffT)e;
                        else a.result->overflow = 1;
                    }
                    if ((ph & 1) == 0) {  // header / plus line: its first byte must be '@' / '+'
                        if (!(first_flags & (ph == 0 ? 2u : 1u))) atomicMax(&a.result->err_pos, ~(unsigned long long)line_start);
                    } else if (ph == 1) {
                        if (kSeq) {
                            if (rix < (uint64_t)a.rec_cap) {
                                a.seq_len[rix] = len;
                                a.gc[rix] = (uint32_t)gsum;
                            } else {
                                a.result->overflow = 1;
                            }
                        }
                    } else if (kQual) {
                        if (rix < (uint64_t)a.rec_cap) {
                            a.qual_len[rix] = len;
                            a.qsum[rix] = ssum - 13 * (int)cr - 33 * (int)len;
                        } else {
                            a.result->overflow = 1;
                        }
                    }
                }
            }
        };
        // 64 records per round: both groups of 32 are loaded before either is worked on (a 4 KiB tile of 150 bp reads has ~45
        // newlines, so this is ONE round of dependent global loads per tile instead of two -- the kernel is latency bound)
        for (int lo = 0; lo < n_events; lo += 64) {
            const int ia = lo + lane, ib = lo + 32 + lane;
            const bool act_a = ia < n_events, act_b = ib < n_events;
            uint2 ra = make_uint2(0u, 0u), rb = make_uint2(0u, 0u);
            if (act_a) ra = a.records[ia < n0 ? off0 + ia : off1 + (ia - n0)];
            if (act_b) rb = a.records[ib < n0 ? off0 + ib : off1 + (ib - n0)];
            int pps_a = __shfl_up_sync(0xffffffffu, (int)ra.x, 1), pps_b = __shfl_up_sync(0xffffffffu, (int)rb.x, 1);
            uint32_t py_a = __shfl_up_sync(0xffffffffu, ra.y, 1), py_b = __shfl_up_sync(0xffffffffu, rb.y, 1);
            const int last_ps_a = __shfl_sync(0xffffffffu, (int)ra.x, 31);
            const uint32_t last_y_a = __shfl_sync(0xffffffffu, ra.y, 31);
            if (lane == 0) {
                pps_a = c_ps;
                py_a = c_y;
                pps_b = last_ps_a;
                py_b = last_y_a;
            }
            emit_line(act_a, ia, ra, pps_a, py_a);
            emit_line(act_b, ib, rb, pps_b, py_b);
            c_ps = __shfl_sync(0xffffffffu, (int)rb.x, 31);
            c_y = __shfl_sync(0xffffffffu, rb.y, 31);
        }
      }
    }
    n_crlf = __reduce_add_sync(0xffffffffu, n_crlf);
    if (lane == 0 && n_crlf) atomicAdd(&a.result->crlf_lines, (unsigned long long)n_crlf);
}

// K2 of the fused flavour: one thread per tile picks the bucket of the tile's true phase and finishes the
// tile's first line; one atomicAdd per warp.
__global__ void __launch_bounds__(256) fastq_fused_combine_kernel(const FastqScanArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = a.n_tiles;
    const int64_t origin = a.begin & ~(int64_t)15;
    const uint64_t init = a.prev ? a.prev->total_lines : 0ull;
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    const FusedPlan plan = make_plan(a);
    __shared__ exb_predicate s_preds[EXB_MAX_PREDICATES];
    if (threadIdx.x < a.n_fused) s_preds[threadIdx.x] = a.fused[threadIdx.x];
    __syncthreads();
    long long cnt = 0, qs = 0, ql = 0;
    bool bad = false;
    for (int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tile < n_tiles; tile += threads) {
        const int n_events = (int)a.tile_cnt[tile];
        const bool is_last = tile == n_tiles - 1;
        if (tile == 0 && a.prev && a.prev->err_pos != 0ull) atomicMax(&a.result->err_pos, a.prev->err_pos);
        if (n_events == 0 && !is_last) continue;
        const uint64_t excl = init + (uint64_t)a.line_base[tile];
        const int64_t tile_base = origin + tile * WT_BYTES;
        const OpenLine open = open_line_before(a.tails, tile, origin, a);
        if (n_events > 0) {
            const FusedTile* ft = a.fused_tiles + tile;
            const int h = (int)(excl & 3);
            const uint32_t cq = ft->cq[h];
            cnt += cq & 0xFFFu;
            ql += cq >> 12;
            qs += ft->qs[h];
            bad = bad || ((ft->bad4 >> h) & 1u) != 0;
            // the tile's first line (it began in an earlier tile, or exactly at this tile's first byte)
            const uint32_t y0 = ft->y0;
            uint32_t len = (uint32_t)(tile_base - open.start) + (uint32_t)rec_pos(y0);
            const uint32_t cr = len > 0 ? rec_cr(y0) : 0u;
            len -= cr;
            if ((h & 1) == 0) {
                if (!(open.flags & (h == 0 ? 2u : 1u))) bad = true;
            } else if (h == 3) {
                const int q1 = ft->ps0 + open.s - 13 * (int)cr - 33 * (int)len;
                // lines longer than a tile can exceed the |sum| < 2^20 precondition of the fast path: use the general test
                FusedPlan p1 = plan;
                if (len > 4096u) p1.simple = 0;
                if (fused_pass(s_preds, a.n_fused, p1, q1, len)) {
                    cnt += 1;
                    qs += q1;
                    ql += len;
                }
            }
        }
        if (is_last) write_final_state(a, excl + (uint64_t)n_events, n_events, tile_base, a.tails[tile], open);
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, dd);
        qs += __shfl_xor_sync(0xffffffffu, qs, dd);
        ql += __shfl_xor_sync(0xffffffffu, ql, dd);
    }
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(a.fused_agg) + 0, (unsigned long long)cnt);
        if (qs) atomicAdd(reinterpret_cast<unsigned long long*>(a.fused_agg) + 3, (unsigned long long)qs);
        if (ql) atomicAdd(reinterpret_cast<unsigned long long*>(a.fused_agg) + 4, (unsigned long long)ql);
        // malformed input: the fused flavour knows THAT a line start contradicts its phase, not where; the host
        // wrapper re-runs the general scan for the offset when it sees this mark (err_pos = n)
        if (any_bad) atomicMax(&a.result->err_pos, ~(unsigned long long)a.n);
    }
}

// K2 of the fused TOTALS flavour (EXB_F_FUSED | EXB_F_SEQ): the bucket of the tile's true phase holds the aggregates of
// its sequence lines 1..; the tile's first line is finished here from the predecessors' tail words (length and G/C
// count of the part before the tile).  fused_agg: [1] sum of lengths, [2] sum of G/C, [3] sequence lines, [5] sum of
// round(gc_content * 2^32)  -- the layout exb_fastq_seq_totals fills from the per-record arrays of the general scan.
__global__ void __launch_bounds__(256) fastq_fused_seq_combine_kernel(const FastqScanArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = a.n_tiles;
    const int64_t origin = a.begin & ~(int64_t)15;
    const uint64_t init = a.prev ? a.prev->total_lines : 0ull;
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    unsigned long long cnt = 0, sl = 0, sg = 0, sf = 0;
    bool bad = false;
    for (int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tile < n_tiles; tile += threads) {
        const int n_events = (int)a.tile_cnt[tile];
        const bool is_last = tile == n_tiles - 1;
        if (tile == 0 && a.prev && a.prev->err_pos != 0ull) atomicMax(&a.result->err_pos, a.prev->err_pos);
        if (n_events == 0 && !is_last) continue;
        const uint64_t excl = init + (uint64_t)a.line_base[tile];
        const int64_t tile_base = origin + tile * WT_BYTES;
        const OpenLine open = open_line_before(a.tails, tile, origin, a);
        if (n_events > 0) {
            const FusedTile* ft = a.fused_tiles + tile;
            const int h = (int)(excl & 3);
            const uint32_t cq = ft->cq[h];
            cnt += cq & 0xFFFu;
            sl += cq >> 12;
            sg += (unsigned long long)(uint32_t)ft->qs[h];
            sf += a.fused_fix[tile * 4 + h];
            bad = bad || ((ft->bad4 >> h) & 1u) != 0;
            const uint32_t y0 = ft->y0;
            uint32_t len = (uint32_t)(tile_base - open.start) + (uint32_t)rec_pos(y0);
            const uint32_t cr = len > 0 ? rec_cr(y0) : 0u;
            len -= cr;
            if ((h & 1) == 0) {
                if (!(open.flags & (h == 0 ? 2u : 1u))) bad = true;
            } else if (h == 1) {
                const uint32_t g1 = (uint32_t)(rec_pg(y0) + open.g);
                cnt += 1;
                sl += len;
                sg += g1;
                sf += gc_fix32(g1, len);
            }
        }
        if (is_last) write_final_state(a, excl + (uint64_t)n_events, n_events, tile_base, a.tails[tile], open);
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, dd);
        sl += __shfl_xor_sync(0xffffffffu, sl, dd);
        sg += __shfl_xor_sync(0xffffffffu, sg, dd);
        sf += __shfl_xor_sync(0xffffffffu, sf, dd);
    }
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        unsigned long long* agg = reinterpret_cast<unsigned long long*>(a.fused_agg);
        if (sl) atomicAdd(agg + 1, sl);
        if (sg) atomicAdd(agg + 2, sg);
        if (cnt) atomicAdd(agg + 3, cnt);
        if (sf) atomicAdd(agg + 5, sf);
        if (any_bad) atomicMax(&a.result->err_pos, ~(unsigned long long)a.n);
    }
}

// =================================================================== sharded COUNT with ONE exchange
// A byte-range shard does not know the phase of its first line until it has heard from its predecessors.  Instead of
// waiting for that (exchange -> compose -> K2 -> reduce: four dependent launches, two of them spinning on NVLink flags),
// K2 is run BEFORE the exchange for all four phases at once: candidate c holds the aggregates of every line of the range
// except its first one under the hypothesis "the range's first line has phase c" (each tile's four buckets are simply
// added to the candidate they belong to), and the first line -- the only one whose length, Phred sum and first byte
// depend on the predecessors -- is described instead of judged.  One 256-byte record per shard is then exchanged and
// EVERY rank combines all records the same way (fastq_combine_records): no second collective.
//   rec[0] lines of the range   rec[1] open_line_start (local)   rec[2] tail_s   rec[3] open-line first-byte flags
//   rec[5] local offset of the range's first newline (-1 = none)  rec[6] its CR flag
//   rec[7] byte sum of the range before that newline              rec[8] first-byte flags of the range
//   rec[9 + 3c ..] count, Phred sum, length sum of candidate c    rec[21] bit c: a line start contradicts candidate c
__global__ void __launch_bounds__(256) fastq_fused_candidates_kernel(const FastqScanArgs a, long long* __restrict__ rec) {
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = a.n_tiles;
    const int64_t origin = a.begin & ~(int64_t)15;
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    const FusedPlan plan = make_plan(a);
    __shared__ exb_predicate s_preds[EXB_MAX_PREDICATES];
    if (threadIdx.x < a.n_fused) s_preds[threadIdx.x] = a.fused[threadIdx.x];
    __syncthreads();
    long long cnt[4] = {0, 0, 0, 0}, qs[4] = {0, 0, 0, 0}, ql[4] = {0, 0, 0, 0};
    uint32_t bad = 0;
    for (int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tile < n_tiles; tile += threads) {
        const int n_events = (int)a.tile_cnt[tile];
        const bool is_last = tile == n_tiles - 1;
        if (tile == 0) rec[8] = tail_byte0_flags(a.tails[0]);
        if (n_events == 0 && !is_last) continue;
        const uint64_t lb = (uint64_t)a.line_base[tile];  // lines of the range before this tile
        const int64_t tile_base = origin + tile * WT_BYTES;
        const OpenLine open = open_line_before(a.tails, tile, origin, a);
        if (n_events > 0) {
            const FusedTile* ft = a.fused_tiles + tile;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int h = (int)((lb + (uint64_t)c) & 3);
                const uint32_t cq = ft->cq[h];
                cnt[c] += cq & 0xFFFu;
                ql[c] += cq >> 12;
                qs[c] += ft->qs[h];
                bad |= ((ft->bad4 >> h) & 1u) << c;
            }
            const uint32_t y0 = ft->y0;
            if (lb == 0) {  // the range's first newline: its line is finished by whoever knows the predecessors
                rec[5] = tile_base + rec_pos(y0);
                rec[6] = rec_cr(y0);
                rec[7] = (long long)ft->ps0 + open.s;
            } else {  // a line that began inside the range
                uint32_t len = (uint32_t)(tile_base - open.start) + (uint32_t)rec_pos(y0);
                const uint32_t cr = len > 0 ? rec_cr(y0) : 0u;
                len -= cr;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int h = (int)((lb + (uint64_t)c) & 3);
                    if ((h & 1) == 0 && !(open.flags & (h == 0 ? 2u : 1u))) bad |= 1u << c;
                }
                const int c3 = (int)((3 - lb) & 3);  // the candidate under which this line is a quality line
                const int q1 = ft->ps0 + open.s - 13 * (int)cr - 33 * (int)len;
                FusedPlan p1 = plan;
                if (len > 4096u) p1.simple = 0;  // see fastq_fused_combine_kernel
                if (fused_pass(s_preds, a.n_fused, p1, q1, len)) {
                    cnt[c3] += 1;
                    qs[c3] += q1;
                    ql[c3] += len;
                }
            }
        }
        if (is_last) {
            write_final_state(a, lb + (uint64_t)n_events, n_events, tile_base, a.tails[tile], open);
            rec[0] = (long long)(lb + (uint64_t)n_events);
            if (n_events > 0) {
                const uint64_t tw = a.tails[tile];
                rec[1] = tile_base + tail_rel_of(tw);
                rec[2] = tail_s_of(tw);
                rec[3] = tail_open_flags(tw);
            } else {
                rec[1] = open.start;
                rec[2] = (long long)open.s + tail_s_of(a.tails[tile]);
                rec[3] = open.flags;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) {
            cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], dd);
            qs[c] += __shfl_xor_sync(0xffffffffu, qs[c], dd);
            ql[c] += __shfl_xor_sync(0xffffffffu, ql[c], dd);
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    // block-level sum before the global atomics: 13 words shared by every warp of the grid would serialise ~9 k
    // same-address atomics each (measured: the kernel took twice as long as K2 with one set of atomics per warp)
    __shared__ long long s_part[8][13];
    const int warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            s_part[warp][3 * c] = cnt[c];
            s_part[warp][3 * c + 1] = qs[c];
            s_part[warp][3 * c + 2] = ql[c];
        }
        s_part[warp][12] = (long long)bad;
    }
    __syncthreads();
    if (threadIdx.x < 13) {
        long long v = 0;
        if (threadIdx.x < 12) {
            for (int w = 0; w < 8; w++) v += s_part[w][threadIdx.x];
        } else {
            for (int w = 0; w < 8; w++) v |= s_part[w][12];
        }
        unsigned long long* r = reinterpret_cast<unsigned long long*>(rec);
        if (v) {
            if (threadIdx.x < 12) atomicAdd(r + 9 + threadIdx.x, (unsigned long long)v);
            else atomicOr(r + 21, (unsigned long long)v);
        }
    }
}

__global__ void fastq_combine_records_kernel(const long long* recs, const int64_t* ranges, int world, FusedPreds fp, long long* total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) fastq_combine_records(recs, ranges, world, fp.p, fp.n, total);
}
cudaError_t fastq_combine_records_launch(const long long* recs, const int64_t* ranges, int world, const exb_predicate* preds, int n_preds,
                                         long long* total, cudaStream_t st) {
    FusedPreds fp;
    fp.n = n_preds;
    for (int i = 0; i < n_preds; i++) fp.p[i] = preds[i];
    fastq_combine_records_kernel<<<1, 32, 0, st>>>(recs, ranges, world, fp, total);
    return cudaGetLastError();
}
cudaError_t fastq_candidates_launch(const FastqScanArgs& a, long long* rec, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(rec, 0, 256, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(rec + 5, 0xFF, 8, st);  // no newline yet
    if (e != cudaSuccess) return e;
    int64_t blocks = (a.n_tiles + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fastq_fused_candidates_kernel<<<dim3((unsigned)blocks), dim3(256), 0, st>>>(a, rec);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ launchers
int64_t fastq_scan_tiles(int64_t begin, int64_t n, int is_final) {
    const int64_t origin = begin & ~(int64_t)15;
    // final range: +1 byte of room for the virtual terminator at n
    const int64_t t = (n + (is_final ? 1 : 0) - origin + WT_BYTES - 1) / WT_BYTES;
    return t > 0 ? t : 1;
}
int64_t fastq_record_slack(int64_t n_tiles) {  // record slots the per-warp bump allocation may leave unused
    const int64_t max_warps = (int64_t)160 * 8 * FQ_WARPS;
    const int64_t warps = n_tiles < max_warps ? n_tiles : max_warps;
    return warps * REC_BLOCK + REC_BLOCK;
}

template <int FLAGS>
static cudaError_t launch_tile_kernel(FastqScanArgs a, cudaStream_t st) {
    constexpr int smem = FqWarpAux<FLAGS>::cta_bytes;
    auto kern = fastq_tile_kernel<FLAGS>;
    static int ctas_per_sm = 0, n_sm = 0;  // per template instance; one device type per process
    cudaError_t e;
    // function attributes belong to a device's context: every device a reader runs on opts in once (a multi-GPU scan
    // launched on device 1 with device 0's settings fails with "invalid argument")
    static std::atomic<unsigned> attr_done{0};
    {
        int cur = 0;
        if ((e = cudaGetDevice(&cur)) != cudaSuccess) return e;
        const unsigned bit = 1u << (cur & 31);
        if (!(attr_done.load() & bit)) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
            attr_done.fetch_or(bit);
        }
    }
    if (ctas_per_sm == 0) {
        int dev = 0, occ = 0, sms = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FQ_THREADS, smem)) != cudaSuccess) return e;
        n_sm = sms;
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    // the kernel keeps tile and row indices in 32 bits: 2^26 tiles = 256 GiB per launch, more than one GPU's HBM
    if (a.n_tiles >= ((int64_t)1 << 26)) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap tm;
    if ((e = make_tensor_map(a.buf, a.begin, a.n, &tm, &a.tma_rows)) != cudaSuccess) return e;
    int64_t grid = (a.n_tiles + FQ_WARPS - 1) / FQ_WARPS;
    const int64_t persistent = (int64_t)n_sm * ctas_per_sm;
    if (grid > persistent) grid = persistent;
    // 0x7F7F7F7F / 0x7B7B7B7B travel as arguments so that ptxas keeps them in registers (see eq_flags)
    kern<<<dim3((unsigned)grid), dim3(FQ_THREADS), smem, st>>>(tm, a, 0x7F7F7F7Fu, 0x7B7B7B7Bu);
    return cudaGetLastError();
}

cudaError_t fastq_tile_launch(const FastqScanArgs& a, int flags, cudaStream_t st) {
    if (flags & EXB_F_FUSED)
        return (flags & EXB_F_SEQ) && !(flags & EXB_F_QUAL) ? launch_tile_kernel<EXB_F_FUSED | EXB_F_SEQ>(a, st) : launch_tile_kernel<EXB_F_FUSED | EXB_F_QUAL>(a, st);
    switch (flags & (EXB_F_SEQ | EXB_F_QUAL)) {
    case 0: return launch_tile_kernel<0>(a, st);
    case EXB_F_SEQ: return launch_tile_kernel<EXB_F_SEQ>(a, st);
    case EXB_F_QUAL: return launch_tile_kernel<EXB_F_QUAL>(a, st);
    default: return launch_tile_kernel<EXB_F_SEQ | EXB_F_QUAL>(a, st);
    }
}

template <typename OffT, int FLAGS>
static cudaError_t launch_emit_kernel(const FastqScanArgs& a, cudaStream_t st) {
    int64_t blocks = (a.n_tiles + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    fastq_emit_kernel<OffT, FLAGS><<<dim3((unsigned)blocks), dim3(256), 0, st>>>(a);
    return cudaGetLastError();
}
template <typename OffT>
static cudaError_t launch_emit(const FastqScanArgs& a, int flags, cudaStream_t st) {
    switch (flags & 7) {
    case 0: return launch_emit_kernel<OffT, 0>(a, st);
    case 1: return launch_emit_kernel<OffT, 1>(a, st);
    case 2: return launch_emit_kernel<OffT, 2>(a, st);
    case 3: return launch_emit_kernel<OffT, 3>(a, st);
    case 4: return launch_emit_kernel<OffT, 4>(a, st);
    case 5: return launch_emit_kernel<OffT, 5>(a, st);
    case 6: return launch_emit_kernel<OffT, 6>(a, st);
    default: return launch_emit_kernel<OffT, 7>(a, st);
    }
}

// The scan's result block (line count of the range, open last line, tail sums) without running K2: everything it holds
// follows from the last tile's directory words.  A byte-range shard exchanges this block BEFORE it knows its phase, so
// the provisional K2 it used to run first (39 us of a 1.45 ms step at C2) was wasted work on every rank but the first.
__global__ void fastq_final_state_kernel(const FastqScanArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t origin = a.begin & ~(int64_t)15;
    const int64_t last = a.n_tiles - 1;
    const uint64_t init = a.prev ? a.prev->total_lines : 0ull;
    const int n_events = (int)a.tile_cnt[last];
    const uint64_t excl = init + (uint64_t)a.line_base[last];
    const OpenLine open = open_line_before(a.tails, last, origin, a);
    write_final_state(a, excl + (uint64_t)n_events, n_events, origin + last * WT_BYTES, a.tails[last], open);
}
cudaError_t fastq_final_state_launch(const FastqScanArgs& a, cudaStream_t st) {
    fastq_final_state_kernel<<<1, 32, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t fastq_emit_launch(const FastqScanArgs& a, int flags, bool wide_offsets, cudaStream_t st) {
    if (flags & EXB_F_FUSED) {
        int64_t blocks = (a.n_tiles + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if ((flags & EXB_F_SEQ) && !(flags & EXB_F_QUAL)) fastq_fused_seq_combine_kernel<<<dim3((unsigned)blocks), dim3(256), 0, st>>>(a);
        else fastq_fused_combine_kernel<<<dim3((unsigned)blocks), dim3(256), 0, st>>>(a);
        return cudaGetLastError();
    }
    return wide_offsets ? launch_emit<uint64_t>(a, flags, st) : launch_emit<uint32_t>(a, flags, st);
}

}  // namespace exb
