// fastq_scan.cu -- single-pass FASTQ line/record scan for sm_100a.
//
// Replaces the record loop of noodles-fastq 0.8 Reader::read_record as driven
// by exon 0.2.6's FASTQ batch reader (call sites: rust/src/arrow_reader.rs:
// 104-118,125-153 in the reference).  Instead of reading a
// record at a time it reads every input byte exactly once.
//
// One CTA (256 threads) owns a tile of SUBS x 16 KiB staged in shared memory
// (cp.async, XOR-swizzled).  Phases:
//
//  A. analysis (byte-parallel, branch-free): thread t owns the 64 contiguous
//     bytes [t*64, t*64+64) of each 16 KiB sub-tile -> 64-bit newline mask,
//     signed byte sum, G/C mask; one packed warp scan + an 8-entry cross-warp
//     scan give every run its tile-level prefix (newlines, byte sum, G/C).
//     Newline positions are scattered, in order, into an event list.
//     The tile's newline count is published at once (CH_AGG).
//  B. event prefixes (event-parallel): one newline per lane computes the byte
//     sum / G,C count of everything before it in the tile.  The part after the
//     last newline becomes the tile's 32-byte "tail record": what a successor
//     needs to finish the line that is open at its start.  It depends on
//     nobody, so it is published without waiting for any other tile.
//  C. chaining: only the newline COUNT is chained (one 64-bit word, status in
//     the top bits, decoupled look-back by warp 0, 64 predecessors per round;
//     the other seven warps sleep at a barrier).
//  D. emission (event-parallel): line length, G/C count (sequence lines) and
//     Phred sum (quality lines) are differences of the prefixes at consecutive
//     newlines; FASTQ's strict 4-line phase (global line index mod 4)
//     disambiguates '@'/'+' inside quality strings.
//
// Outputs are single-writer stores (no atomics except the rare error path):
//   line_end[g]            position of the newline ending line g      (F_LINES)
//   seq_len[r], gc[r]      per record                                 (F_SEQ)
//   qual_len[r], qsum[r]   per record; qsum = sum((signed char)c - 33) (F_QUAL)
// HBM traffic: input once + 4..16 B per line of results.
#include "common.cuh"
#include "exon_b200_internal.h"

namespace exb {

constexpr int SUB_BYTES = TILE_BYTES;  // 16 KiB analysed per pass of the 256 threads

struct alignas(16) TailRec {  // the part of a tile after its last newline (the whole tile if it has none)
    int64_t line_start;       // absolute offset of the byte after the tile's last newline
    int64_t tail_s;           // signed byte sum of that part
    int64_t tail_g;           // G/C count of that part
    uint64_t state;           // 0 = not published yet, 1 = no newline in the tile, 2 = line_start valid
};

// State of the line that is open at the start of `tile`: walk the predecessors' tail
// records back to the one that holds the line's start (usually tile-1).
__device__ __noinline__ void open_line_before(const TailRec* recs, int64_t tile, const FastqScanArgs& a, int64_t& start, int64_t& ts,
                                              int64_t& tg) {
    ts = 0;
    tg = 0;
    for (int64_t k = tile - 1; k >= 0; k--) {
        uint64_t st;
        do {
            st = ld_acquire_u64(&recs[k].state);
        } while (st == 0);
        const uint4 v0 = ld_cg_u4(reinterpret_cast<const uint4*>(&recs[k]));
        const uint4 v1 = ld_cg_u4(reinterpret_cast<const uint4*>(&recs[k]) + 1);
        ts += (int64_t)(((uint64_t)v0.w << 32) | v0.z);
        tg += (int64_t)(((uint64_t)v1.y << 32) | v1.x);
        if (st == 2) {
            start = (int64_t)(((uint64_t)v0.y << 32) | v0.x);
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

// 16-bit equality mask of a 16-byte chunk, bits in byte order.  The 0x80 flags of two
// words are folded into one byte by IDP.4A with weights 1,2,4,8 / 16,32,64,128 (the
// products carry a factor 128 that one shift removes): 4 IDP + 2 ops instead of 12.
__device__ __forceinline__ uint32_t flags_to_mask16(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    uint32_t lo = __dp4a(m0, 0x08040201u, 0u);
    lo = __dp4a(m1, 0x80402010u, lo);
    uint32_t hi = __dp4a(m2, 0x08040201u, 0u);
    hi = __dp4a(m3, 0x80402010u, hi);
    return (lo >> 7) | (hi << 1);
}
__device__ __forceinline__ uint32_t nl_mask16(const uint4& v) {
    return flags_to_mask16(eq_bytes(v.x, 0x0A0A0A0Au), eq_bytes(v.y, 0x0A0A0A0Au), eq_bytes(v.z, 0x0A0A0A0Au), eq_bytes(v.w, 0x0A0A0A0Au));
}
__device__ __forceinline__ uint32_t gc_mask16b(const uint4& v) {
    return flags_to_mask16(gc_bytes(v.x), gc_bytes(v.y), gc_bytes(v.z), gc_bytes(v.w));
}

template <int FLAGS, int SUBS>
struct FqSmem {
    static constexpr bool kSeq = (FLAGS & EXB_F_SEQ) != 0, kQual = (FLAGS & EXB_F_QUAL) != 0;
    static constexpr int EV_CAP = 512 * SUBS;
    static constexpr int off_data = 0;
    static constexpr int off_sexcl = off_data + SUBS * SUB_BYTES;
    static constexpr int off_gexcl = off_sexcl + (kQual ? SUBS * BLOCK_THREADS * 4 : 0);
    static constexpr int off_gm = off_gexcl + (kSeq ? SUBS * BLOCK_THREADS * 4 : 0);
    static constexpr int off_evps = off_gm + (kSeq ? SUBS * BLOCK_THREADS * 8 : 0);
    static constexpr int off_evpg = off_evps + (kQual ? EV_CAP * 4 : 0);
    static constexpr int off_evpos = off_evpg + (kSeq ? EV_CAP * 4 : 0);
    static constexpr int total = off_evpos + EV_CAP * 2;
};

template <typename OffT, int FLAGS, int SUBS>
__global__ void __launch_bounds__(BLOCK_THREADS) fastq_scan_kernel(FastqScanArgs a) {
    using SM = FqSmem<FLAGS, SUBS>;
    constexpr bool kLines = (FLAGS & EXB_F_LINES) != 0;
    constexpr bool kSeq = SM::kSeq, kQual = SM::kQual;
    constexpr int TILE = SUBS * SUB_BYTES;
    constexpr int EV_CAP = SM::EV_CAP;

    extern __shared__ __align__(16) uint8_t smem[];
    uint4* s_data = reinterpret_cast<uint4*>(smem + SM::off_data);
    int* s_sexcl = reinterpret_cast<int*>(smem + SM::off_sexcl);
    int* s_gexcl = reinterpret_cast<int*>(smem + SM::off_gexcl);
    uint64_t* s_gm = reinterpret_cast<uint64_t*>(smem + SM::off_gm);
    int* ev_ps = reinterpret_cast<int*>(smem + SM::off_evps);
    int* ev_pg = reinterpret_cast<int*>(smem + SM::off_evpg);
    uint16_t* ev_pos = reinterpret_cast<uint16_t*>(smem + SM::off_evpos);
    __shared__ int s_wt[SUBS][3][WARPS];  // per sub-tile: warp totals of (newlines, byte sum, G/C)
    __shared__ int64_t s_tile_id;
    __shared__ uint64_t s_excl;
    __shared__ int s_batch_carry[3];  // last event of the previous window: pos, ps (incl. newline), pg

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint8_t* __restrict__ buf = a.buf;
    const int64_t origin = a.begin & ~(int64_t)15;
    uint64_t* chain = reinterpret_cast<uint64_t*>(a.slots);
    TailRec* recs = reinterpret_cast<TailRec*>(chain + ((a.n_tiles + 1) & ~(int64_t)1));

    if (t == 0) s_tile_id = (int64_t)atomicAdd(a.ticket, 1ull);
    __syncthreads();
    const int64_t tile = s_tile_id;
    const int64_t tile_base = origin + tile * TILE;

    // ---- staging: interior tiles take the cheap path (one 64-bit base, 32-bit offsets)
    if (tile_base >= a.begin && tile_base + TILE <= a.n) {
        const uint8_t* src = buf + tile_base + t * 16;
#pragma unroll
        for (int s = 0; s < SUBS; s++)
#pragma unroll
            for (int i = 0; i < RUN_CHUNKS; i++)
                cp_async16(&s_data[s * TILE_CHUNKS + swz(i * BLOCK_THREADS + t)], src + (s * SUB_BYTES + i * BLOCK_THREADS * 16), 16);
        cp_async_commit();
    } else {
#pragma unroll 1
        for (int s = 0; s < SUBS; s++) stage_tile(s_data + s * TILE_CHUNKS, buf, tile_base + (int64_t)s * SUB_BYTES, origin, a.n);
    }
    cp_async_wait<0>();
    __syncthreads();

    uint8_t* sbytes = reinterpret_cast<uint8_t*>(s_data);
    // tile-local byte index -> shared memory byte (the swizzle permutes chunks inside a sub-tile only)
    auto sidx = [](int li) -> int { return ((li >> 14) << 14) + swz((li >> 4) & (TILE_CHUNKS - 1)) * 16 + (li & 15); };

    // Rare per-launch patches (uniform per CTA): bytes before `begin` in the first
    // chunk are filler; an unterminated last line gets a virtual '\n' at index n.
    const bool has_begin_pad = (tile == 0 && a.begin != origin);
    const bool has_eof = a.is_final && (a.n >= tile_base && a.n < tile_base + TILE);
    if (has_begin_pad || has_eof) {
        if (t == 0) {
            if (has_begin_pad)
                for (int64_t i = origin; i < a.begin; i++) sbytes[sidx((int)(i - origin))] = 0;
            if (has_eof && a.n > a.begin && buf[a.n - 1] != '\n') sbytes[sidx((int)(a.n - tile_base))] = '\n';
        }
        __syncthreads();
    }
    auto any_byte = [&](int64_t abs_pos) -> int {  // byte at an absolute offset < tile end that may lie before this tile
        int64_t li = abs_pos - tile_base;
        if (li >= 0) return sbytes[sidx((int)li)];
        return (abs_pos >= (a.prev ? 0 : a.begin)) ? (int)buf[abs_pos] : -1;
    };

    // ---- A. analysis of the SUBS sub-tiles; scatters the events with rank in [win_lo, win_lo + EV_CAP)
    int total_cnt = 0, total_s = 0, total_g = 0;
    auto analyse = [&](int win_lo) {
        total_cnt = total_s = total_g = 0;
#pragma unroll 1
        for (int s = 0; s < SUBS; s++) {
            const uint4* d = s_data + s * TILE_CHUNKS;
            uint64_t pm, gm = 0;
            int rs = 0;  // signed byte sum of the run
            {
                const uint4 c0 = d[swz(4 * t + 0)], c1 = d[swz(4 * t + 1)], c2 = d[swz(4 * t + 2)], c3 = d[swz(4 * t + 3)];
                pm = ((uint64_t)(nl_mask16(c2) | (nl_mask16(c3) << 16)) << 32) | (nl_mask16(c0) | (nl_mask16(c1) << 16));
                if (kSeq) gm = ((uint64_t)(gc_mask16b(c2) | (gc_mask16b(c3) << 16)) << 32) | (gc_mask16b(c0) | (gc_mask16b(c1) << 16));
                if (kQual) rs = sbyte_sum16(c3, sbyte_sum16(c2, sbyte_sum16(c1, sbyte_sum16(c0, 0))));
            }
            const int cnt = __popcll(pm);
            const int gtot = kSeq ? __popcll(gm) : 0;
            // warp scan: newlines (<= 2048 per warp: 12 bits) and byte sum (|.| <= 2^18: 20 bits signed) share one word
            const uint32_t packed = ((uint32_t)cnt << 20) + (uint32_t)rs;
            const uint32_t incl = warp_incl_scan_u32(packed);
            const uint32_t ex = incl - packed;
            const int ex_s = ((int)(ex << 12)) >> 12;
            const int ex_cnt = (int)((ex - (uint32_t)ex_s) >> 20);
            int ex_g = 0, in_g = 0;
            if (kSeq) {
                in_g = (int)warp_incl_scan_u32((uint32_t)gtot);
                ex_g = in_g - gtot;
            }
            if (lane == 31) {
                const int in_s = ((int)(incl << 12)) >> 12;
                s_wt[s][0][warp] = (int)((incl - (uint32_t)in_s) >> 20);
                s_wt[s][1][warp] = in_s;
                s_wt[s][2][warp] = in_g;
            }
            __syncthreads();
            // cross-warp exclusive scan of the 8 warp totals (every warp redoes it in its low lanes)
            int wc = lane < WARPS ? s_wt[s][0][lane] : 0, ws = lane < WARPS ? s_wt[s][1][lane] : 0;
            int wg = (kSeq && lane < WARPS) ? s_wt[s][2][lane] : 0;
            int ic = wc, is_ = ws, ig = wg;
#pragma unroll
            for (int dd = 1; dd < WARPS; dd <<= 1) {
                const int oc = __shfl_up_sync(0xffffffffu, ic, dd), os = __shfl_up_sync(0xffffffffu, is_, dd);
                const int og = kSeq ? __shfl_up_sync(0xffffffffu, ig, dd) : 0;
                if (lane >= dd) {
                    ic += oc;
                    is_ += os;
                    ig += og;
                }
            }
            const int off_cnt = __shfl_sync(0xffffffffu, ic - wc, warp), off_s = __shfl_sync(0xffffffffu, is_ - ws, warp);
            const int off_g = kSeq ? __shfl_sync(0xffffffffu, ig - wg, warp) : 0;
            const int sc = __shfl_sync(0xffffffffu, ic, WARPS - 1), ss = __shfl_sync(0xffffffffu, is_, WARPS - 1);
            const int sg = kSeq ? __shfl_sync(0xffffffffu, ig, WARPS - 1) : 0;
            const int run = s * BLOCK_THREADS + t;
            if (kQual) s_sexcl[run] = total_s + off_s + ex_s;
            if (kSeq) {
                s_gexcl[run] = total_g + off_g + ex_g;
                s_gm[run] = gm;
            }
            {  // events of this run, in order
                int rank = total_cnt + off_cnt + ex_cnt - win_lo;
                uint64_t m = pm;
                const int p0 = s * SUB_BYTES + t * RUN_BYTES;
                while (m) {
                    const int k = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    if ((unsigned)rank < (unsigned)EV_CAP) ev_pos[rank] = (uint16_t)(p0 + k);
                    rank++;
                }
            }
            total_cnt += sc;
            total_s += ss;
            total_g += sg;
        }
        __syncthreads();  // event list complete
    };
    // ---- B. prefix sums at the newlines of the current window (one per lane)
    auto event_prefixes = [&](int nev) {
        if (kQual || kSeq) {
            for (int i = t; i < nev; i += BLOCK_THREADS) {
                const int pos = ev_pos[i];
                const int run = pos >> 6, k = pos & 63;
                if (kQual) {
                    const uint4* d = s_data + (run >> 8) * TILE_CHUNKS;
                    const int q = 4 * (run & 255);
                    int ps = s_sexcl[run];
                    const int j = k >> 4;
                    if (j > 0) ps = sbyte_sum16(d[swz(q)], ps);
                    if (j > 1) ps = sbyte_sum16(d[swz(q + 1)], ps);
                    if (j > 2) ps = sbyte_sum16(d[swz(q + 2)], ps);
                    ev_ps[i] = ps + sbyte_sum_prefix16(d[swz(q + j)], k & 15);
                }
                if (kSeq) ev_pg[i] = s_gexcl[run] + __popcll(s_gm[run] & low_bits64(k));
            }
        }
        __syncthreads();
    };
    // ---- D. emission of the current window
    auto emit = [&](int lo, int nev, uint64_t excl) {
        for (int i = t; i < nev; i += BLOCK_THREADS) {
            const uint64_t g = excl + (uint64_t)(lo + i);  // global index of the line this newline ends
            if (g >= a.max_lines) continue;
            const int pos = ev_pos[i];
            const int64_t e = tile_base + pos;
            long long len, ssum = 0, gsum = 0;
            int first_byte;  // first byte of the line
            if (i > 0 || lo > 0) {
                const int ppos = i > 0 ? (int)ev_pos[i - 1] : s_batch_carry[0];
                len = pos - ppos - 1;
                if (kQual) ssum = ev_ps[i] - (i > 0 ? ev_ps[i - 1] + 10 : s_batch_carry[1]);
                if (kSeq) gsum = ev_pg[i] - (i > 0 ? ev_pg[i - 1] : s_batch_carry[2]);
                first_byte = sbytes[sidx(ppos + 1)];
            } else {  // the line that was open when the tile began
                int64_t st0, s0, g0;
                open_line_before(recs, tile, a, st0, s0, g0);
                len = e - st0;
                if (kQual) ssum = ev_ps[i] + s0;
                if (kSeq) gsum = ev_pg[i] + g0;
                first_byte = any_byte(st0);
            }
            // a CR directly before a real LF is stripped; the virtual '\n' at EOF strips nothing
            int cr = 0;
            if (len > 0 && !(a.is_final && e == a.n)) cr = (pos > 0 ? (int)sbytes[sidx(pos - 1)] : any_byte(e - 1)) == '\r';
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
                if (first_byte != (ph == 0 ? '@' : '+')) atomicMin(&a.result->err_pos, (unsigned long long)(e - len - cr));
            } else if (r < (uint64_t)a.rec_cap) {
                if (ph == 1) {
                    if (kSeq) {
                        a.seq_len[r] = (uint32_t)len;
                        a.gc[r] = (uint32_t)gsum;
                    }
                } else if (kQual) {
                    a.qual_len[r] = (uint32_t)len;
                    a.qsum[r] = (int32_t)(ssum - 13 * cr - 33 * len);
                }
            } else {
                a.result->overflow = 1;
            }
        }
    };

    analyse(0);
    const int n_events = total_cnt;
    const uint64_t init = a.prev ? a.prev->total_lines : 0ull;
    if (t == 0) {  // the count is all the chain needs: publish it before anything else
        if (tile == 0) st_release_u64(&chain[0], CH_INC | ((init + (uint64_t)n_events) & CH_VAL));
        else st_release_u64(&chain[tile], CH_AGG | (uint64_t)n_events);
        if (a.prev && tile == 0) {  // chained range: errors of earlier ranges stay visible in the last result
            if (a.prev->err_pos != ~0ull) atomicMin(&a.result->err_pos, a.prev->err_pos);
            if (a.prev->overflow) a.result->overflow = 1;
        }
    }
    const bool dense = n_events > EV_CAP;  // more newlines than the event window holds (rare): windows are re-analysed
    if (dense) analyse(((n_events - 1) / EV_CAP) * EV_CAP);
    const int last_nev = n_events - ((n_events - 1) / EV_CAP) * EV_CAP;  // events in the window that holds the last one
    event_prefixes(dense ? last_nev : n_events);

    // tail record (local information only), then the look-back of the count by warp 0
    TailRec mine;
    if (n_events > 0) {
        const int li = (dense ? last_nev : n_events) - 1;
        mine.state = 2;
        mine.line_start = tile_base + ev_pos[li] + 1;
        mine.tail_s = kQual ? total_s - (ev_ps[li] + 10) : 0;
        mine.tail_g = kSeq ? total_g - ev_pg[li] : 0;
    } else {
        mine.state = 1;
        mine.line_start = 0;
        mine.tail_s = total_s;
        mine.tail_g = total_g;
    }
    if (warp == 0) {
        if (lane == 0) {
            recs[tile].line_start = mine.line_start;
            recs[tile].tail_s = mine.tail_s;
            recs[tile].tail_g = mine.tail_g;
            __threadfence();
            st_release_u64(&recs[tile].state, mine.state);
        }
        const uint64_t ex = warp_lookback(chain, tile, (uint64_t)n_events, init);
        if (lane == 0) s_excl = ex;
    }
    __syncthreads();
    const uint64_t excl = s_excl;

    if (tile == a.n_tiles - 1 && t == 0) {  // final state of this range (chaining / host)
        int64_t st = mine.line_start, s2 = mine.tail_s, g2 = mine.tail_g;
        if (mine.state != 2) {
            open_line_before(recs, tile, a, st, s2, g2);
            s2 += mine.tail_s;
            g2 += mine.tail_g;
        }
        a.result->total_lines = excl + (uint64_t)n_events;
        a.result->open_line_start = st;
        a.result->tail_s = s2;
        a.result->tail_g = g2;
    }

    if (!dense) {
        emit(0, n_events, excl);
    } else {
        for (int lo = 0; lo < n_events; lo += EV_CAP) {
            const int nev = min(EV_CAP, n_events - lo);
            __syncthreads();
            analyse(lo);
            event_prefixes(nev);
            emit(lo, nev, excl);
            __syncthreads();
            if (t == 0) {  // carry the window's last event into the next one
                s_batch_carry[0] = ev_pos[nev - 1];
                s_batch_carry[1] = kQual ? ev_ps[nev - 1] + 10 : 0;
                s_batch_carry[2] = kSeq ? ev_pg[nev - 1] : 0;
            }
        }
    }
}

// ------------------------------------------------------------------ launcher
template <typename OffT, int FLAGS, int SUBS>
static cudaError_t launch_one(FastqScanArgs a, cudaStream_t st) {
    constexpr int smem = FqSmem<FLAGS, SUBS>::total;
    auto kern = fastq_scan_kernel<OffT, FLAGS, SUBS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    a.n_tiles = (a.n_tiles + SUBS - 1) / SUBS;  // the caller counted 16 KiB sub-tiles
    kern<<<dim3((unsigned)a.n_tiles), dim3(BLOCK_THREADS), smem, st>>>(a);
    return cudaGetLastError();
}

template <typename OffT>
static cudaError_t launch_fastq(const FastqScanArgs& a, int flags, cudaStream_t st) {
    constexpr int SUBS = EXB_FASTQ_SUBS;
    switch (flags & 7) {
    case 0: return launch_one<OffT, 0, SUBS>(a, st);
    case 1: return launch_one<OffT, 1, SUBS>(a, st);
    case 2: return launch_one<OffT, 2, SUBS>(a, st);
    case 3: return launch_one<OffT, 3, SUBS>(a, st);
    case 4: return launch_one<OffT, 4, SUBS>(a, st);
    case 5: return launch_one<OffT, 5, SUBS>(a, st);
    case 6: return launch_one<OffT, 6, SUBS>(a, st);
    default: return launch_one<OffT, 7, SUBS>(a, st);
    }
}

cudaError_t fastq_scan_launch(const FastqScanArgs& a, int flags, bool wide_offsets, cudaStream_t st) {
    return wide_offsets ? launch_fastq<uint64_t>(a, flags, st) : launch_fastq<uint32_t>(a, flags, st);
}

}  // namespace exb
