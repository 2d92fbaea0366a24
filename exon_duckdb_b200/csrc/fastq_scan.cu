// fastq_scan.cu -- single-pass FASTQ line/record scan for sm_100a.
//
// Replaces the record loop of noodles-fastq 0.8 Reader::read_record as driven
// by exon 0.2.6's FASTQ batch reader (call sites: rust/src/arrow_reader.rs:
// 104-118,125-153 in the reference).  Instead of reading a
// record at a time it reads every input byte exactly once:
//
//   tile (16 KiB) -> shared memory (cp.async, swizzled)
//   per thread: 64 contiguous bytes -> newline bitmask, signed byte sum, G/C mask
//   warp scan + cross-warp totals   -> tile aggregate = number of newlines
//   block-wide decoupled look-back  -> global line index of every newline
//   per newline: line length, G/C count (sequence lines), Phred sum (quality
//   lines) as differences of prefix sums; FASTQ's strict 4-line phase
//   (line index mod 4) disambiguates '@'/'+' inside quality strings.
//
// Only the newline COUNT is chained between tiles (one 64-bit word, status in
// the top bits).  What a tile needs to finish the line that was open at its
// start -- where that line began and its partial sums -- is purely local to
// the predecessor that holds the line's start, so each tile publishes a
// 32-byte "tail record" with no dependency on anybody and the one thread that
// owns the tile's first newline reads its predecessor's record directly.
//
// Outputs are single-writer stores (no atomics except the rare error path):
//   line_end[g]            position of the newline ending line g      (F_LINES)
//   seq_len[r], gc[r]      per record                                 (F_SEQ)
//   qual_len[r], qsum[r]   per record; qsum = sum((signed char)c - 33) (F_QUAL)
// HBM traffic: input once + 4..16 B per line of results.
#include "common.cuh"
#include "exon_b200_internal.h"

namespace exb {

struct alignas(16) TailRec {  // the part of a tile after its last newline (the whole tile if it has none)
    int64_t line_start;       // absolute offset of the byte after the tile's last newline
    int64_t tail_s;           // signed byte sum of that part
    int64_t tail_g;           // G/C count of that part
    uint64_t has_nl;
};

struct WarpLast {  // where the line open at the end of a warp's 2 KiB started
    int start_local;   // tile-local index of the byte after the warp's last newline
    int a_s, a_g;      // warp-relative prefix sums up to and including that newline
    int valid;
};

// packed (count, gc, signed sum) in one 64-bit lane value: 21 bits each
__device__ __forceinline__ uint64_t pack3(int cnt, int g, int s) {
    return ((uint64_t)cnt << 42) + ((uint64_t)g << 21) + (uint64_t)(int64_t)s;
}
__device__ __forceinline__ void unpack3(uint64_t v, int& cnt, int& g, int& s) {
    int64_t sv = ((int64_t)(v << 43)) >> 43;  // sign-extend low 21 bits
    uint64_t rest = (v - (uint64_t)sv) >> 21;
    s = (int)sv;
    g = (int)(rest & 0x1FFFFFu);
    cnt = (int)(rest >> 21);
}

__device__ __forceinline__ TailRec ld_rec(const TailRec* p) {
    TailRec r;
    uint4 a = ld_cg_u4(reinterpret_cast<const uint4*>(p));
    uint4 b = ld_cg_u4(reinterpret_cast<const uint4*>(p) + 1);
    r.line_start = (int64_t)(((uint64_t)a.y << 32) | a.x);
    r.tail_s = (int64_t)(((uint64_t)a.w << 32) | a.z);
    r.tail_g = (int64_t)(((uint64_t)b.y << 32) | b.x);
    r.has_nl = ((uint64_t)b.w << 32) | b.z;
    return r;
}

// State of the line that is open at the start of `tile`: walk the predecessors' tail
// records back to the one that holds the line's start (usually tile-1).
__device__ __forceinline__ void open_line_before(const TailRec* recs, int64_t tile, const FastqScanArgs& a, int64_t& start, int64_t& ts,
                                                 int64_t& tg) {
    ts = 0;
    tg = 0;
    for (int64_t k = tile - 1; k >= 0; k--) {
        TailRec r = ld_rec(&recs[k]);
        ts += r.tail_s;
        tg += r.tail_g;
        if (r.has_nl) {
            start = r.line_start;
            return;
        }
    }
    if (a.prev) {
        start = a.prev->open_line_start;
        ts += a.prev->tail_s;
        tg += a.prev->tail_g;
    } else {
        start = a.begin;
    }
}

template <typename OffT, int FLAGS>
__global__ void __launch_bounds__(BLOCK_THREADS) fastq_scan_kernel(FastqScanArgs a) {
    constexpr bool kLines = (FLAGS & EXB_F_LINES) != 0;
    constexpr bool kSeq = (FLAGS & EXB_F_SEQ) != 0;
    constexpr bool kQual = (FLAGS & EXB_F_QUAL) != 0;

    __shared__ uint4 s_tile[TILE_CHUNKS];
    __shared__ uint64_t s_warp_tot[WARPS];
    __shared__ WarpLast s_warp_last[WARPS];
    __shared__ LookbackSmem<1> s_lb;
    __shared__ int64_t s_tile_id;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint8_t* __restrict__ buf = a.buf;
    const int64_t origin = a.begin & ~(int64_t)15;
    uint64_t* chain = reinterpret_cast<uint64_t*>(a.slots);
    TailRec* recs = reinterpret_cast<TailRec*>(chain + ((a.n_tiles + 1) & ~(int64_t)1));

    if (t == 0) s_tile_id = (int64_t)atomicAdd(a.ticket, 1ull);
    __syncthreads();
    const int64_t tile = s_tile_id;
    const int64_t tile_base = origin + tile * TILE_BYTES;

    stage_tile(s_tile, buf, tile_base, origin, a.n);
    cp_async_wait<0>();
    __syncthreads();

    // Rare per-launch patches (uniform per CTA): bytes before `begin` in the first
    // chunk are filler; an unterminated last line gets a virtual '\n' at index n.
    const bool has_begin_pad = (tile == 0 && a.begin != origin);
    const bool has_eof = a.is_final && (a.n >= tile_base && a.n < tile_base + TILE_BYTES);
    if (has_begin_pad || has_eof) {
        if (t == 0) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(s_tile);
            if (has_begin_pad)
                for (int64_t i = origin; i < a.begin; i++) sb[swz(0) * 16 + (int)(i - origin)] = 0;
            if (has_eof && a.n > a.begin && buf[a.n - 1] != '\n') {
                int li = (int)(a.n - tile_base);
                sb[swz(li >> 4) * 16 + (li & 15)] = '\n';
            }
        }
        __syncthreads();
    }
    const uint8_t* sbytes = reinterpret_cast<const uint8_t*>(s_tile);
    auto smem_byte = [&](int li) -> int { return sbytes[swz(li >> 4) * 16 + (li & 15)]; };
    auto any_byte = [&](int64_t abs_pos) -> int {  // byte at an absolute offset that may lie before this tile
        int64_t li = abs_pos - tile_base;
        if (li >= 0) return smem_byte((int)li);
        return (abs_pos >= (a.prev ? 0 : a.begin)) ? (int)buf[abs_pos] : -1;
    };

    // ---- per-thread analysis of its 64-byte run
    uint64_t pm = 0, gm = 0;
    int cs0 = 0, cs1 = 0, cs2 = 0, cs3 = 0;
    {
        uint4 c0 = s_tile[swz(4 * t + 0)], c1 = s_tile[swz(4 * t + 1)];
        uint4 c2 = s_tile[swz(4 * t + 2)], c3 = s_tile[swz(4 * t + 3)];
        uint32_t lo = eq_mask16(c0, 0x0A0A0A0Au) | (eq_mask16(c1, 0x0A0A0A0Au) << 16);
        uint32_t hi = eq_mask16(c2, 0x0A0A0A0Au) | (eq_mask16(c3, 0x0A0A0A0Au) << 16);
        pm = ((uint64_t)hi << 32) | lo;
        if (kSeq) {
            uint32_t glo = gc_mask16(c0) | (gc_mask16(c1) << 16);
            uint32_t ghi = gc_mask16(c2) | (gc_mask16(c3) << 16);
            gm = ((uint64_t)ghi << 32) | glo;
        }
        if (kQual) {
            cs0 = sbyte_sum16(c0, 0);
            cs1 = sbyte_sum16(c1, cs0);
            cs2 = sbyte_sum16(c2, cs1);
            cs3 = sbyte_sum16(c3, cs2);
        }
    }
    const int cnt = __popcll(pm);
    const int gtot = kSeq ? __popcll(gm) : 0;

    // prefix sums inside the run: signed byte sum / G/C count of run bytes [0, k)
    auto run_prefix_s = [&](int k) -> int {
        if (!kQual) return 0;
        int j = k >> 4;
        int base = j == 0 ? 0 : (j == 1 ? cs0 : (j == 2 ? cs1 : cs2));
        uint4 ch = s_tile[swz(4 * t + j)];
        return base + sbyte_sum_prefix16(ch, k & 15);
    };
    auto run_prefix_g = [&](int k) -> int { return kSeq ? __popcll(gm & low_bits64(k)) : 0; };

    // ---- warp scan of (newlines, G/C, byte sum)
    const uint64_t packed = pack3(cnt, gtot, cs3);
    const uint64_t incl = warp_incl_scan_u64(packed);
    if (lane == 31) s_warp_tot[warp] = incl;
    int ex_cnt, ex_g, ex_s;  // warp-relative exclusive prefix of this thread
    unpack3(incl - packed, ex_cnt, ex_g, ex_s);

    // line-start info after this thread's last newline (warp-relative sums)
    const uint32_t has = __ballot_sync(0xffffffffu, cnt > 0);
    int my_start = 0, my_as = 0, my_ag = 0;
    if (cnt > 0) {
        int k = 63 - __clzll((long long)pm);
        my_start = t * RUN_BYTES + k + 1;
        my_as = ex_s + run_prefix_s(k) + (kQual ? 10 : 0);
        my_ag = ex_g + run_prefix_g(k);
    }
    if (has == 0) {
        if (lane == 0) s_warp_last[warp].valid = 0;
    } else if (lane == 31 - __clz(has)) {
        s_warp_last[warp] = WarpLast{my_start, my_as, my_ag, 1};  // sums still warp-relative
    }
    __syncthreads();

    // ---- tile totals, tail record, look-back of the newline count
    int off_cnt = 0, off_g = 0, off_s = 0;  // totals of the preceding warps
    int tc = 0, tg = 0, ts = 0;             // tile totals
    int lw = -1, lw_s = 0, lw_g = 0;        // last warp with a newline and the totals before it
    for (int w = 0; w < WARPS; w++) {
        if (w == warp) {
            off_cnt = tc;
            off_g = tg;
            off_s = ts;
        }
        if (s_warp_last[w].valid) {
            lw = w;
            lw_s = ts;
            lw_g = tg;
        }
        int c_, g_, s_;
        unpack3(s_warp_tot[w], c_, g_, s_);
        tc += c_;
        tg += g_;
        ts += s_;
    }
    TailRec mine;
    mine.has_nl = lw >= 0;
    if (lw >= 0) {
        mine.line_start = tile_base + s_warp_last[lw].start_local;
        mine.tail_s = ts - (lw_s + s_warp_last[lw].a_s);
        mine.tail_g = tg - (lw_g + s_warp_last[lw].a_g);
    } else {
        mine.line_start = 0;
        mine.tail_s = ts;
        mine.tail_g = tg;
    }
    if (t == 0) {  // published by block_lookback's release (thread 0 stores, fences, then writes the chain word)
        recs[tile] = mine;
        if (a.prev && tile == 0) {  // chained range: errors of earlier ranges stay visible in the last result
            if (a.prev->err_pos != ~0ull) atomicMin(&a.result->err_pos, a.prev->err_pos);
            if (a.prev->overflow) a.result->overflow = 1;
        }
    }
    const uint64_t agg[1] = {(uint64_t)tc};
    const uint64_t init[1] = {a.prev ? a.prev->total_lines : 0ull};
    uint64_t excl[1];
    block_lookback<1>(chain, tile, agg, init, excl, &s_lb);

    if (tile == a.n_tiles - 1 && t == 0) {  // final state of this range (chaining / host)
        int64_t st, s2, g2;
        if (mine.has_nl) {
            st = mine.line_start;
            s2 = mine.tail_s;
            g2 = mine.tail_g;
        } else {
            open_line_before(recs, tile, a, st, s2, g2);
            s2 += mine.tail_s;
            g2 += mine.tail_g;
        }
        a.result->total_lines = excl[0] + (uint64_t)tc;
        a.result->open_line_start = st;
        a.result->tail_s = s2;
        a.result->tail_g = g2;
    }
    if (cnt == 0) return;

    // ---- per-newline emission
    const int tile_ex_s = off_s + ex_s, tile_ex_g = off_g + ex_g;
    uint64_t g = excl[0] + (uint64_t)(off_cnt + ex_cnt);  // global index of my first line end

    int64_t cur_start, cur_as, cur_ag;  // the line open at the start of my run
    {
        const uint32_t before = has & ((1u << lane) - 1u);
        const int src = before ? 31 - __clz(before) : lane;
        // lanes without newlines returned above: `has` is exactly the mask of live lanes
        int st = __shfl_sync(has, my_start, src);
        int as_ = __shfl_sync(has, my_as, src);
        int ag_ = __shfl_sync(has, my_ag, src);
        if (before) {
            cur_start = tile_base + st;
            cur_as = off_s + as_;
            cur_ag = off_g + ag_;
        } else {
            int w = warp - 1;
            while (w >= 0 && !s_warp_last[w].valid) w--;
            if (w >= 0) {
                int ws = 0, wg = 0;
                for (int i = 0; i < w; i++) {
                    int c_, g_, s_;
                    unpack3(s_warp_tot[i], c_, g_, s_);
                    ws += s_;
                    wg += g_;
                }
                cur_start = tile_base + s_warp_last[w].start_local;
                cur_as = ws + s_warp_last[w].a_s;
                cur_ag = wg + s_warp_last[w].a_g;
            } else {  // my first newline closes the line that was open when the tile began
                int64_t st0, s0, g0;
                open_line_before(recs, tile, a, st0, s0, g0);
                cur_start = st0;
                cur_as = -s0;
                cur_ag = -g0;
            }
        }
    }

    const int run0 = t * RUN_BYTES;
    const int64_t run_base = tile_base + run0;
    uint64_t m = pm;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int64_t e = run_base + k;  // absolute position of this newline
        const int64_t ps = tile_ex_s + run_prefix_s(k);
        const int64_t pg = tile_ex_g + run_prefix_g(k);
        if (g < a.max_lines) {
            int64_t len = e - cur_start;
            // a CR directly before a real LF is stripped; the virtual '\n' at EOF strips nothing
            int cr = 0;
            if (len > 0 && !(a.is_final && e == a.n)) cr = (run0 + k > 0 ? smem_byte(run0 + k - 1) : any_byte(e - 1)) == '\r';
            len -= cr;
            const int ph = (int)(g & 3);
            const uint64_t r = g >> 2;
            if (kLines) {
                if (g < (uint64_t)a.line_cap)
                    reinterpret_cast<OffT*>(a.line_end)[g] = (OffT)e;
                else
                    a.result->overflow = 1;
            }
            if ((ph & 1) == 0) {  // header / plus line: its first byte must be '@' / '+'
                if (any_byte(cur_start) != (ph == 0 ? '@' : '+')) atomicMin(&a.result->err_pos, (unsigned long long)cur_start);
            } else if (r < (uint64_t)a.rec_cap) {
                if (ph == 1) {
                    if (kSeq) {
                        a.seq_len[r] = (uint32_t)len;
                        a.gc[r] = (uint32_t)(pg - cur_ag);
                    }
                } else {
                    if (kQual) {
                        a.qual_len[r] = (uint32_t)len;
                        a.qsum[r] = (int32_t)((ps - cur_as) - 13 * cr - 33 * len);
                    }
                }
            } else {
                a.result->overflow = 1;
            }
        }
        cur_start = e + 1;
        cur_as = ps + (kQual ? 10 : 0);
        cur_ag = pg;
        g++;
    }
}

// ------------------------------------------------------------------ launcher
template <typename OffT>
static cudaError_t launch_fastq(const FastqScanArgs& a, int flags, cudaStream_t st) {
    dim3 grid((unsigned)a.n_tiles), block(BLOCK_THREADS);
#define EXB_CASE(F)                                                   \
    case F:                                                           \
        fastq_scan_kernel<OffT, F><<<grid, block, 0, st>>>(a);        \
        break;
    switch (flags & 7) {
        EXB_CASE(0)
        EXB_CASE(1)
        EXB_CASE(2)
        EXB_CASE(3)
        EXB_CASE(4)
        EXB_CASE(5)
        EXB_CASE(6)
        EXB_CASE(7)
    }
#undef EXB_CASE
    return cudaGetLastError();
}

cudaError_t fastq_scan_launch(const FastqScanArgs& a, int flags, bool wide_offsets, cudaStream_t st) {
    return wide_offsets ? launch_fastq<uint64_t>(a, flags, st) : launch_fastq<uint32_t>(a, flags, st);
}

}  // namespace exb
