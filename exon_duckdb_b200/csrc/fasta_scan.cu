// fasta_scan.cu -- FASTA record scan + sequence compaction for sm_100a (v2: chain-free byte pass).
//
// Replaces noodles-fasta 0.27 Reader::read_definition / read_sequence as driven
// by exon 0.2.6's FASTA batch reader (reference call sites: rust/src/
// arrow_reader.rs:104-118,125-153).  Semantics:
//   * a record starts at a line whose FIRST byte is '>' ('>' elsewhere is data);
//   * the sequence is every byte of the following non-header lines, with each
//     line's LF (and a CR directly before that LF) removed; blank lines add nothing.
//
// What a byte MEANS depends on one bit of context -- "is the line that is open at this
// point a header line?" -- and its output position on three running sums (records, kept
// sequence bytes, G/C among them).  v1 chained those across tiles inside one kernel
// (decoupled look-back) and was latency bound (350 GB/s on B200).  v2 splits the work
// where the dependency is, exactly like the FASTQ scan:
//
//  K1  fasta_tile_kernel   every input byte once, NO inter-tile dependency.  Same skeleton
//      as fastq_tile_kernel (tma_tile.cuh): a warp owns a 4 KiB tile fetched by ONE TMA
//      box load into 128B-swizzled shared memory, lane = 128-byte row, dense newline and
//      G/C masks (SWAR + IDP.4A), then one newline per lane: the byte after it ('>' = a
//      record starts), the byte before it (CR), the G/C prefix at it.  A tile's summary
//      is 16 bytes: records started, kept bytes / G,C of the lines that begin inside the
//      tile, kept bytes / G,C of the HEAD piece (the part of the line that was already open
//      at the tile's first byte -- counted only if that line is not a header), and how the
//      tile ends (at a line start / inside a header / inside a sequence line / no newline).
//  --  a scan over the tile summaries (3 tiny kernels: reduce, spine, downsweep).  The carried
//      state is {inside header, inside sequence line, at a line start}; a tile is a function
//      on it, and function composition is the scan operator, so newline-free tiles (a
//      250 Mbp unwrapped contig is 61 k of them) cost nothing extra.
//  K2  fasta_emit_kernel   only tiles that start a record or end a header line are read
//      again (C3: 6000 of 745 k tiles) to write hdr_start / hdr_end / seq_off / gc_prefix.
//  K3  fasta_compact_kernel (only when the sequence column is wanted) second pass over the
//      bytes, TMA fed, each line copied to its final position through a shared-memory
//      staging row so the global stores are 16-byte coalesced.
//
// HBM traffic: gc_content per contig (no sequence column) = input once + 41 B per tile;
// with the sequence column = input twice + output once.
#include <cuda.h>

#include <atomic>

#include "common.cuh"
#include "exon_b200_internal.h"
#include "tma_tile.cuh"

namespace exb {

constexpr int FA_WARPS = 4;
constexpr int FA_THREADS = FA_WARPS * 32;
constexpr int FA_EV_CAP = 64;

// carried state: what is open at a byte position
constexpr uint32_t ST_H = 0, ST_S = 1, ST_L = 2;  // inside a header line / inside a sequence line / at a line start

// ---------------------------------------------------------------- tile summary (16 bytes)
// a: [12:0] records started after a newline inside the tile | [13] has a newline | [14] the line open after the last
//    newline is a header | [15] the last valid byte is a newline | [16] the first valid byte is '>' | [17] no valid byte
// rest: kept | gc << 16   (lines that begin inside the tile)        head: kept | gc << 16   (the carried-in piece)
struct alignas(16) FaTile {
    uint32_t a, rest, head, pos0;  // pos0: [15:0] position of the first newline (hdr_end of a carried-in header), 0xFFFF = none;
                                   //       [31:16] FT_REGULAR tiles with a newline: position of the last one
};
constexpr uint32_t FT_HAS_NL = 1u << 13, FT_HDR_LAST = 1u << 14, FT_ENDS_NL = 1u << 15, FT_B0_GT = 1u << 16, FT_EMPTY = 1u << 17;
constexpr uint32_t FT_REGULAR = 1u << 18;  // a full interior tile whose only bytes below 0x40 are LFs (K1 tells K3: bit 3 of tile_state)

__device__ __forceinline__ uint32_t ft_eff(uint32_t a, uint32_t s) { return s == ST_L ? ((a & FT_B0_GT) ? ST_H : ST_S) : s; }
__device__ __forceinline__ uint32_t ft_out(uint32_t a, uint32_t s) {
    if (a & FT_EMPTY) return s;
    if (a & FT_ENDS_NL) return ST_L;
    if (a & FT_HAS_NL) return (a & FT_HDR_LAST) ? ST_H : ST_S;
    return ft_eff(a, s);
}
__device__ __forceinline__ uint32_t ft_nhs(uint32_t a, uint32_t s) { return (a & 0x1FFFu) + ((s == ST_L && (a & FT_B0_GT) && !(a & FT_EMPTY)) ? 1u : 0u); }

// aggregate of a run of tiles as a function of the state at its first byte
struct FaAgg {
    uint32_t out[3];
    uint64_t n[3], k[3], g[3];
};
__device__ __forceinline__ FaAgg fa_identity() {
    FaAgg r;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        r.out[s] = s;
        r.n[s] = r.k[s] = r.g[s] = 0;
    }
    return r;
}
__device__ __forceinline__ FaAgg fa_of_tile(const FaTile& t) {
    FaAgg r;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const bool head_counts = !(t.a & FT_EMPTY) && ft_eff(t.a, s) == ST_S;
        r.out[s] = ft_out(t.a, s);
        r.n[s] = ft_nhs(t.a, s);
        r.k[s] = (t.rest & 0xFFFFu) + (head_counts ? (t.head & 0xFFFFu) : 0u);
        r.g[s] = (t.rest >> 16) + (head_counts ? (t.head >> 16) : 0u);
    }
    return r;
}
__device__ __forceinline__ FaAgg fa_compose(const FaAgg& a, const FaAgg& b) {  // a first, then b
    FaAgg r;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const uint32_t m = a.out[s];
        r.out[s] = m == 0 ? b.out[0] : (m == 1 ? b.out[1] : b.out[2]);
        r.n[s] = a.n[s] + (m == 0 ? b.n[0] : (m == 1 ? b.n[1] : b.n[2]));
        r.k[s] = a.k[s] + (m == 0 ? b.k[0] : (m == 1 ? b.k[1] : b.k[2]));
        r.g[s] = a.g[s] + (m == 0 ? b.g[0] : (m == 1 ? b.g[1] : b.g[2]));
    }
    return r;
}
struct FaVal {  // the scan's running value: state + sums at a tile's first byte
    uint32_t st;
    uint64_t n, k, g;
};
__device__ __forceinline__ FaVal fa_apply(const FaVal& v, const FaAgg& a) {
    FaVal r;
    const uint32_t s = v.st;
    r.st = s == 0 ? a.out[0] : (s == 1 ? a.out[1] : a.out[2]);
    r.n = v.n + (s == 0 ? a.n[0] : (s == 1 ? a.n[1] : a.n[2]));
    r.k = v.k + (s == 0 ? a.k[0] : (s == 1 ? a.k[1] : a.k[2]));
    r.g = v.g + (s == 0 ? a.g[0] : (s == 1 ? a.g[1] : a.g[2]));
    return r;
}

// ---------------------------------------------------------------- shared memory of one warp
struct FaAux {
    static constexpr int off_gm = 0;                      // u64[64]: G/C mask of each 64-byte half row
    static constexpr int off_gex = off_gm + 64 * 8;       // int[64]: G/C prefix at each half row
    static constexpr int off_ev = off_gex + 64 * 4;       // u16[EV_CAP + 2]: newline positions (+ a dump slot)
    static constexpr int off_bar = off_ev + FA_EV_CAP * 2 + 16;  // 2 mbarriers
    static constexpr int total = off_bar + 16;
};
template <bool kCompact>
struct FaSmem {
    static constexpr int out_bytes = kCompact ? WT_BYTES + 32 : 0;  // staging row of the compacted bytes
    static constexpr int per_warp = 2 * WT_BYTES;
    static constexpr int aux0 = FA_WARPS * per_warp;
    static constexpr int out0 = aux0 + FA_WARPS * FaAux::total;
    static constexpr int cta_bytes = ((out0 + 15) & ~15) + FA_WARPS * out_bytes;
};

// geometry of one tile inside the range [begin, n)
struct FaGeom {
    int s0;        // first valid position (begin - origin in tile 0, else 0)
    int data_end;  // positions [s0, data_end) hold bytes of the input
    bool virt;     // position data_end holds the virtual '\n' of an unterminated last line
};
__device__ __forceinline__ FaGeom fa_geom(const FastaScanArgs& a, int64_t origin, int64_t tile) {
    FaGeom g;
    const int64_t tile_base = origin + tile * WT_BYTES;
    g.s0 = tile == 0 ? (int)(a.begin - origin) : 0;
    const int64_t left = a.n - tile_base;
    g.data_end = left < 0 ? 0 : (left > WT_BYTES ? WT_BYTES : (int)left);
    if (g.data_end < g.s0) g.data_end = g.s0;
    g.virt = a.is_final && a.n > a.begin && a.n >= tile_base && a.n < tile_base + WT_BYTES && a.buf[a.n - 1] != '\n';
    return g;
}

// Edge fixes after the tile's bytes have arrived (uniform per warp, rare): rows the tensor map does not cover, bytes
// before `begin`, the virtual terminator.  `have_rows` = rows of this tile that came through TMA (or were staged).
__device__ __forceinline__ void fa_fix_edges(const FastaScanArgs& a, uint8_t* sbytes, int64_t origin, int64_t tile, int64_t full_rows, int lane,
                                             const FaGeom& g) {
    const int64_t row0 = tile * WT_ROWS;
    const int64_t tile_base = origin + tile * WT_BYTES;
    const bool partial = row0 + WT_ROWS > full_rows;
    const bool has_begin_pad = (tile == 0 && a.begin != origin);
    if (!(partial || has_begin_pad || g.virt)) return;
    if (row0 >= full_rows) {  // nothing came through TMA: clear the buffer
        for (int i = lane; i < WT_BYTES / 16; i += 32) reinterpret_cast<uint4*>(sbytes)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }
    if (partial) {  // bytes of the last, incomplete row (TMA zero-filled everything beyond the tensor)
        const int64_t tail0 = origin + full_rows * ROW_BYTES;
        const int64_t lo = tail0 > tile_base ? tail0 : tile_base;
        for (int64_t p = lo + lane; p < a.n && p < tile_base + WT_BYTES; p += 32) sbytes[sidx((int)(p - tile_base))] = a.buf[p];
    }
    __syncwarp();
    if (lane == 0) {
        if (has_begin_pad)
            for (int i = 0; i < g.s0; i++) sbytes[sidx(i)] = 0;
        if (g.virt) sbytes[sidx(g.data_end)] = '\n';
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes before this buffer's next TMA fill
    __syncwarp();
}

// Copy `len` bytes from tile position `src` (128B-swizzled buffer) to staging offset `dst`: dst-aligned 4-byte words in
// the body (two aligned source words + funnel shift), bytes at the ragged ends.  Work items first, first + step, ...
// so that one lane (0, 1) or a whole warp (lane, 32) can run the same line.
__device__ __forceinline__ void fa_copy(uint8_t* s_out, const uint8_t* sbytes, int dst, int src, int len, int first, int step) {
    int head = (4 - (dst & 3)) & 3;
    if (head > len) head = len;
    for (int i = first; i < head; i += step) s_out[dst + i] = sbytes[sidx(src + i)];
    dst += head;
    src += head;
    len -= head;
    const int nw = len >> 2;
    const int bs = (src & 3) * 8, sa = src & ~3;
    uint32_t* ow = reinterpret_cast<uint32_t*>(s_out + dst);
    for (int k = first; k < nw; k += step) {
        const int p0 = sa + 4 * k;
        const int p1 = p0 + 4 < WT_BYTES ? p0 + 4 : p0;  // bs == 0 whenever the clamp matters (the line ends at the tile's end)
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(sbytes + sidx(p0));
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(sbytes + sidx(p1));
        ow[k] = __funnelshift_r(w0, w1, bs);
    }
    const int done = nw * 4;
    for (int i = done + first; i < len; i += step) s_out[dst + i] = sbytes[sidx(src + i)];
}

// shared-memory accessors on 32-bit shared-window addresses
__device__ __forceinline__ uint32_t fa_lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 fa_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void fa_sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t fa_lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
// tile-local byte index -> byte offset in the 128B-swizzled tile buffer (sidx as two instructions)
__device__ __forceinline__ uint32_t fa_swz(uint32_t li) { return li ^ ((li >> 3) & 0x70u); }
__device__ __forceinline__ void fa_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void fa_sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory"); }

// inclusive warp scan; the shuffle's predicate output (source lane in range) guards the add: two instructions per step
__device__ __forceinline__ uint32_t fa_incl_scan_u32p(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n .reg .pred p;\n .reg .u32 t;\n shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n @p add.u32 %0, %0, t;\n}\n" : "+r"(v) : "r"(d));
    return v;
}
// 0xFF in the low k bytes of a word (k <= 0: none, k >= 4: all)
__device__ __forceinline__ uint32_t bytes_below32c(int k) { return k >= 4 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << (8 * k)) - 1u)); }

// 128 x (number of bytes < 0x40 among the 16): bit 7 of x | x << 1 is set iff bit 7 or bit 6 of the byte is
__device__ __forceinline__ uint32_t low_count128(const uint4& v, uint32_t acc) {
    acc = __dp4a(~(v.x | (v.x << 1)) & 0x80808080u, 0x01010101u, acc);
    acc = __dp4a(~(v.y | (v.y << 1)) & 0x80808080u, 0x01010101u, acc);
    acc = __dp4a(~(v.z | (v.z << 1)) & 0x80808080u, 0x01010101u, acc);
    acc = __dp4a(~(v.w | (v.w << 1)) & 0x80808080u, 0x01010101u, acc);
    return acc;
}

// ---------------------------------------------------------------- the tile analysis (shared by K1 / K2 / K3)
// MODE 0: summary.  MODE 1: per-record outputs (state and bases known).  MODE 2: compaction (state and kept base known).
struct FaTileIn {  // what K2 / K3 know about the tile from the scan
    uint32_t st;
    uint32_t regular;  // K3: K1 found only LFs below 0x40 in this (full, interior) tile
    int64_t rec_base, kept_base, gc_base;
};

template <int MODE>
__device__ __forceinline__ void fa_tile(const FastaScanArgs& a, const uint8_t* sbytes, uint8_t* aux, int64_t origin, int64_t tile, const FaGeom& g,
                                        uint32_t c7f, uint32_t c7b, const FaTileIn& in, FaTile* summary, uint8_t* s_out) {
    const int lane = threadIdx.x & 31;
    uint64_t* s_gm = reinterpret_cast<uint64_t*>(aux + FaAux::off_gm);
    int* s_gex = reinterpret_cast<int*>(aux + FaAux::off_gex);
    uint16_t* ev_pos = reinterpret_cast<uint16_t*>(aux + FaAux::off_ev);
    const uint4* d = reinterpret_cast<const uint4*>(sbytes);
    const int64_t tile_base = origin + tile * WT_BYTES;
    const uint32_t pat_nl = c7f & 0x0A0A0A0Au, pat_gc = c7f & 0x43434343u;

    // ---- K3, evenly wrapped tile: no dense pass at all.  K1 found only LFs below 0x40 in this tile (tile_state bit 3) and
    // left their number and the positions of the first and the last one in its summary.  If the LFs sit on the lattice
    // n0 + k P (lines of W = P - 1 bases: what every FASTA writer produces) -- checked by probing the lattice points, two
    // or three bytes a lane: n_events LFs on n_events lattice points are all of them -- the position of the o-th kept
    // byte has a closed form, o + (LFs before it), and the pass turns OUTPUT-centric: a lane builds whole 16-byte chunks
    // of the compacted column (one unaligned 16-byte read of the tile: five swizzled words + funnel shifts; the bytes
    // behind an LF inside the chunk moved up by one) and stores them straight to global memory, coalesced.  No masks, no
    // per-row piece loops, no staging row (the staged path below executed ~1650 warp instructions a tile:
    // profiles/r02e_fasta_compact_ncu.txt).
    if (MODE == 2) {
        if (in.regular != 0 && tile != 0 && g.s0 == 0 && g.data_end == WT_BYTES && !g.virt && in.st != ST_H) {
            const FaTile ft = *summary;  // K1's summary of this tile
            const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sbytes);
            int n0 = WT_BYTES, P = 1 << 20, n_events = 0;
            bool even = true;
            if ((ft.pos0 & 0xFFFFu) != 0xFFFFu) {
                n0 = (int)(ft.pos0 & 0xFFFFu);
                const int c_last = (int)(ft.pos0 >> 16);
                n_events = WT_BYTES - n0 - (int)(ft.rest & 0xFFFFu);
                if (n_events >= 2) {
                    P = (c_last - n0) / (n_events - 1);
                    even = P * (n_events - 1) == c_last - n0 && P >= 17;
                }
                if (even) {
                    bool ok = true;
                    for (int k = lane; k < n_events; k += 32) ok = ok && fa_lds8(sb + fa_swz((uint32_t)(n0 + k * P))) == (uint32_t)'\n';
                    even = __all_sync(0xffffffffu, ok);
                }
            }
            if (even) {
                const int total = WT_BYTES - n_events;
                const int64_t G0 = in.kept_base;
                if (G0 + total > a.seq_cap) {
                    if (lane == 0) a.result->overflow = 1;
                    return;
                }
                const int W = P - 1;
                const uint32_t magic = 0xFFFFFFFFu / (uint32_t)W + 1u;  // floor(x / W) = umulhi(x, magic) for x W < 2^32
                // byte masks "first k bytes of a 16-byte chunk" for the chunks an LF falls into: a table in the warp's G/C
                // scratch (unused on this path; rewritten per tile because the general path overwrites it)
                const uint32_t tab = (uint32_t)__cvta_generic_to_shared(s_gm);
                if (lane < 16)
                    fa_sts128(tab + 16u * (uint32_t)lane, bytes_below32c(lane), bytes_below32c(lane - 4), bytes_below32c(lane - 8), bytes_below32c(lane - 12));
                __syncwarp();
                uint8_t* __restrict__ gout = a.seq_out + G0;  // kept byte o of the tile -> gout[o]
                auto lf_before = [&](int o) -> int { return o < n0 ? 0 : 1 + (int)__umulhi((uint32_t)(o - n0), magic); };
                int h = (int)((16 - (G0 & 15)) & 15);  // kept bytes in front of the first whole 16-byte chunk
                h = h < total ? h : total;
                const int n_chunks = (total - h) >> 4;
                if (lane < h) gout[lane] = (uint8_t)fa_lds8(sb + fa_swz((uint32_t)(lane + lf_before(lane))));
                const int t0 = h + 16 * n_chunks + lane;  // ... and behind the last one (fewer than 16)
                if (t0 < total) gout[t0] = (uint8_t)fa_lds8(sb + fa_swz((uint32_t)(t0 + lf_before(t0))));
                for (int c = lane; c < n_chunks; c += 32) {
                    const int o = h + 16 * c;
                    const int m0 = lf_before(o), m1 = lf_before(o + 15);
                    const int i0 = o + m0;
                    const int bs = (i0 & 3) * 8;
                    const uint32_t q = (uint32_t)(i0 & ~3);
                    const uint32_t w0 = fa_lds32(sb + fa_swz(q)), w1 = fa_lds32(sb + fa_swz(q + 4)), w2 = fa_lds32(sb + fa_swz(q + 8)),
                                   w3 = fa_lds32(sb + fa_swz(q + 12)), w4 = fa_lds32(sb + fa_swz(q + 16));
                    uint4 v = make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs),
                                         __funnelshift_r(w3, w4, bs));
                    if (m1 != m0) {  // W >= 16: one LF at most; chunk bytes [0, k) stay, [k, 16) come from one byte further on
                        const int k = n0 + (m1 - 1) * W - o;
                        const uint32_t nx = w4 >> bs;  // source byte i0 + 16 in its low byte
                        const uint4 u = make_uint4(__funnelshift_r(v.x, v.y, 8), __funnelshift_r(v.y, v.z, 8), __funnelshift_r(v.z, v.w, 8),
                                                   __funnelshift_r(v.w, nx, 8));
                        const uint4 km = fa_lds128(tab + 16u * (uint32_t)k);
                        v.x = (v.x & km.x) | (u.x & ~km.x);
                        v.y = (v.y & km.y) | (u.y & ~km.y);
                        v.z = (v.z & km.z) | (u.z & ~km.z);
                        v.w = (v.w & km.w) | (u.w & ~km.w);
                    }
                    *reinterpret_cast<uint4*>(gout + o) = v;
                }
                __syncwarp();  // every lane is done with the tile buffer (and the mask table) before the next TMA fill
                return;
            }
        }
    }
    // ---- A. dense masks of the lane's row
    uint64_t pm[2], gm[2];
    uint32_t low128 = 0;  // 128 x the number of bytes < 0x40 in the row (MODE 0 / 2: anything but LF among them = irregular tile)
    {
        const int sw = lane & 7;
        const uint4* row = d + lane * 8;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 c0 = row[(4 * h + 0) ^ sw], c1 = row[(4 * h + 1) ^ sw], c2 = row[(4 * h + 2) ^ sw], c3 = row[(4 * h + 3) ^ sw];
            if (MODE == 0) {  // (the compaction pass reads K1's verdict from tile_state instead of counting again)
                low128 = low_count128(c0, low128);
                low128 = low_count128(c1, low128);
                low128 = low_count128(c2, low128);
                low128 = low_count128(c3, low128);
            }
            pm[h] = ((uint64_t)nl_mask32r(c2, c3, c7f, pat_nl) << 32) | nl_mask32r(c0, c1, c7f, pat_nl);
            if (MODE != 2)  // the compaction pass moves bytes: no G/C bookkeeping
                gm[h] = ((uint64_t)gc_mask32r(c2, c3, c7b, c7f, pat_gc) << 32) | gc_mask32r(c0, c1, c7b, c7f, pat_gc);
            else
                gm[h] = 0;
        }
    }
    const int cnt = __popcll(pm[0]) + __popcll(pm[1]);
    const int g0 = __popcll(gm[0]), g1 = __popcll(gm[1]);
    const uint32_t packed = ((uint32_t)cnt << 16) + (uint32_t)(g0 + g1);
    const uint32_t incl = fa_incl_scan_u32p(packed);
    const uint32_t ex = incl - packed;
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    const int ex_cnt = (int)(ex >> 16), n_events = (int)(tot >> 16);
    const int ex_g = (int)(ex & 0xFFFFu), total_g = (int)(tot & 0xFFFFu);

    // ---- fast path of the summary pass: a full interior tile whose only bytes below 0x40 are its LFs has no '>' (no
    // record starts, no header), no CR and no padding, so its summary follows from the dense masks alone -- the
    // per-newline walk below (2-3 rounds for 60-column lines) is skipped.  Header tiles carry digits / spaces / '>' and
    // take the general path.  C3: 739 k of 745 k tiles are regular.
    if (MODE == 0) {
        const bool regular = tile != 0 && g.s0 == 0 && g.data_end == WT_BYTES && !g.virt;
        const bool odd = (int)(low128 >> 7) != cnt;
        if (regular && !__any_sync(0xffffffffu, odd)) {
            const uint32_t have = __ballot_sync(0xffffffffu, cnt > 0);
            FaTile t;
            if (have == 0) {  // one piece of a long line
                t.a = FT_REGULAR;
                t.rest = 0;
                t.head = (uint32_t)WT_BYTES | ((uint32_t)total_g << 16);
                t.pos0 = 0xFFFFu;
            } else {
                const int first = __ffs((int)have) - 1, last = 31 - __clz((int)have);
                const int fr = pm[0] ? __ffsll((long long)pm[0]) - 1 : 64 + __ffsll((long long)pm[1]) - 1;       // first LF of my row
                const int lr = pm[1] ? 127 - __clzll((long long)pm[1]) : 63 - __clzll((long long)pm[0]);        // last LF of my row
                const int gb = ex_g + (fr < 64 ? __popcll(gm[0] & low_bits64(fr)) : g0 + __popcll(gm[1] & low_bits64(fr - 64)));
                const int pos0 = __shfl_sync(0xffffffffu, lane * ROW_BYTES + fr, first);
                const int head_g = __shfl_sync(0xffffffffu, gb, first);
                const int c_last = __shfl_sync(0xffffffffu, lane * ROW_BYTES + lr, last);
                t.a = FT_REGULAR | FT_HAS_NL | (c_last + 1 >= WT_BYTES ? FT_ENDS_NL : 0u);
                t.rest = (uint32_t)(WT_BYTES - pos0 - n_events) | ((uint32_t)(total_g - head_g) << 16);
                t.head = (uint32_t)pos0 | ((uint32_t)head_g << 16);
                t.pos0 = (uint32_t)pos0 | ((uint32_t)c_last << 16);  // + the last LF: K3 derives the line pitch from the two
            }
            if (lane == 0) *summary = t;
            return;
        }
    }
    // ---- fast path of the compaction pass: the same regular tile, entered in a sequence line (or at a line start: no
    // '>' in the tile, so that line is sequence too).  Every byte but the LFs is kept, so the lane copies the (up to
    // four) LF-free pieces of ITS OWN row to the staging row at lane * 128 - (LFs before the row): no newline walk, no
    // scans, and source words come from one swizzled row (address math is 3 ops a word).
    bool fast2 = false;
    if (MODE == 2) {
        fast2 = in.regular != 0 && tile != 0 && g.s0 == 0 && g.data_end == WT_BYTES && !g.virt && in.st != ST_H;
        if (fast2) {
            // 32-bit shared-window addresses + ld/st.shared: no generic-pointer arithmetic in the copy loop (with generic
            // pointers the word loop cost ~30 instructions a word: ncu r02b)
            const uint32_t row_u32 = (uint32_t)__cvta_generic_to_shared(sbytes) + (uint32_t)(lane * ROW_BYTES);
            const uint32_t out_u32 = (uint32_t)__cvta_generic_to_shared(s_out);
            // word w of the lane's row: the 128B swizzle XORs the 16-byte chunk index with (row & 7), i.e. bits 2..4 of w
            const int x4 = (lane & 7) << 2;
            auto src_word = [&](int w) -> uint32_t { return fa_lds32(row_u32 + (uint32_t)((w ^ x4) << 2)); };
            int dst = (int)(in.kept_base & 15) + lane * ROW_BYTES - ex_cnt;
            int start = 0;
            uint64_t m0 = pm[0], m1 = pm[1];
            while (start < ROW_BYTES) {
                int end;
                if (m0) {
                    end = __ffsll((long long)m0) - 1;
                    m0 &= m0 - 1;
                } else if (m1) {
                    end = 64 + __ffsll((long long)m1) - 1;
                    m1 &= m1 - 1;
                } else {
                    end = ROW_BYTES;
                }
                int len = end - start, src = start;
                // bytes up to the first 4-byte boundary of the destination
                while (len > 0 && (dst & 3)) {
                    fa_sts8(out_u32 + (uint32_t)dst++, src_word(src >> 2) >> ((src & 3) * 8));
                    src++;
                    len--;
                }
                const int nw = len >> 2, bs = (src & 3) * 8, sw = src >> 2;
                const uint32_t ow = out_u32 + (uint32_t)dst;
                if (nw > 0) {
                    // ONE loop for aligned and unaligned sources (the lanes of a warp mix both: two loops ran one after
                    // the other).  bs != 0: the last byte copied lies in word sw + nw, so every word read exists; bs == 0:
                    // the funnel shift returns w0 and the look-ahead word, clamped into the row, is never used.
                    uint32_t w0 = src_word(sw);
                    for (int k = 0; k < nw; k++) {
                        const int nx = sw + k + 1;
                        const uint32_t w1 = src_word(nx < 31 ? nx : 31);
                        fa_sts32(ow + 4u * (uint32_t)k, __funnelshift_r(w0, w1, bs));
                        w0 = w1;
                    }
                }
                dst += nw * 4;
                src += nw * 4;
                len -= nw * 4;
                while (len > 0) {
                    fa_sts8(out_u32 + (uint32_t)dst++, src_word(src >> 2) >> ((src & 3) * 8));
                    src++;
                    len--;
                }
                start = end + 1;
            }
        }
    }
    if (!fast2) {  // only the newline walk reads these
        s_gm[2 * lane] = gm[0];
        s_gm[2 * lane + 1] = gm[1];
        s_gex[2 * lane] = ex_g;
        s_gex[2 * lane + 1] = ex_g + g0;
    }

    auto scatter = [&](int win_lo) {
        int rank = ex_cnt - win_lo;
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            uint32_t m = (uint32_t)(pm[wd >> 1] >> ((wd & 1) * 32));
            const int base = lane * ROW_BYTES + wd * 32;
            while (m) {
                if ((unsigned)rank < (unsigned)FA_EV_CAP) ev_pos[rank] = (uint16_t)(base + __ffs((int)m) - 1);
                m &= m - 1;
                rank++;
            }
        }
    };
    if (!fast2) scatter(0);
    __syncwarp();
    auto gc_before = [&](int pos) -> int { return s_gex[pos >> 6] + __popcll(s_gm[pos >> 6] & low_bits64(pos & 63)); };

    // what the tile's first valid byte says
    const int b0 = g.data_end > g.s0 ? (int)sbytes[sidx(g.s0)] : 0;
    const bool b0_gt = b0 == '>';
    uint32_t eff = ST_S;
    if (MODE != 0) eff = ft_eff(b0_gt ? FT_B0_GT : 0u, in.st);

    // ---- B. one newline per lane
    uint32_t f_nhs = 0, f_kept = 0, f_gc = 0;  // lines 1.. (sequence lines only)
    uint32_t head_kept = 0, head_gc = 0, pos0 = 0xFFFFu;
    int c_pos = g.s0 - 1, c_pg = 0;  // last newline of the previous pass (virtual one before the first valid byte)
    uint32_t c_hdr = 0;               // the line after it is a header
    // MODE 1/2: running totals before the current pass (exclusive over lines)
    int64_t run_kept = 0, run_gc = 0, run_rec = 0;
    if (MODE != 0) {
        run_rec = in.rec_base + ((in.st == ST_L && b0_gt && g.data_end > g.s0) ? 1 : 0);
        run_kept = in.kept_base;
        run_gc = in.gc_base;
        if (MODE == 1 && in.st == ST_L && b0_gt && g.data_end > g.s0 && lane == 0) {  // a record starts with the tile's first byte
            const int64_t r = in.rec_base;
            if (r < a.rec_cap) {
                a.hdr_start[r] = tile_base + g.s0;
                a.seq_off[r] = in.kept_base;
                a.gc_prefix[r] = in.gc_base;
            } else {
                a.result->overflow = 1;
            }
        }
    }
    const int64_t cr_limit = a.is_final ? a.n : (a.halo_n > a.n ? a.halo_n : a.n);  // a CR is dropped iff a REAL LF follows it
    int pos = 0, pg = 0;
    uint32_t hdr_after = 0;
    for (int lo = 0; lo < (fast2 ? 0 : n_events); lo += 32) {
        if (lo > 0 && (lo & (FA_EV_CAP - 1)) == 0) {
            __syncwarp();
            scatter(lo);
            __syncwarp();
        }
        const bool active = lo + lane < n_events;
        uint32_t cr = 0;
        if (active) {
            pos = ev_pos[(lo & (FA_EV_CAP - 1)) + lane];
            pg = gc_before(pos);
            const int before = pos > 0 ? (int)sbytes[sidx(pos - 1)] : 0;
            const int after = pos + 1 < WT_BYTES ? (int)sbytes[sidx(pos + 1)] : 0;
            const bool is_virtual = g.virt && pos == g.data_end;
            cr = (pos - 1 >= g.s0 && before == '\r' && !is_virtual) ? 1u : 0u;
            hdr_after = (pos + 1 < g.data_end && after == '>') ? 1u : 0u;
        } else {
            hdr_after = 0;
        }
        int ppos = __shfl_up_sync(0xffffffffu, pos, 1), ppg = __shfl_up_sync(0xffffffffu, pg, 1);
        uint32_t phdr = __shfl_up_sync(0xffffffffu, hdr_after, 1);
        if (lane == 0) {
            ppos = c_pos;
            ppg = c_pg;
            phdr = c_hdr;
        }
        const bool is_head = (lo + lane == 0);
        uint32_t len = 0, lgc = 0;
        bool line_is_hdr = false;
        if (active) {
            len = (uint32_t)(pos - ppos - 1) - cr;
            lgc = (uint32_t)(pg - ppg);
            line_is_hdr = is_head ? (eff == ST_H) : (phdr != 0);
            if (is_head) {
                head_kept = len;
                head_gc = lgc;
                pos0 = (uint32_t)pos;
            }
        }
        if (MODE == 0) {
            if (active && !is_head && !line_is_hdr) {
                f_kept += len;
                f_gc += lgc;
            }
            f_nhs += hdr_after;
        } else {
            // exclusive prefixes over the lines of this pass
            const uint32_t ck = (active && !line_is_hdr) ? len : 0u, cg = (active && !line_is_hdr) ? lgc : 0u;
            const uint32_t pk = (ck << 13) | hdr_after;  // kept (<= 4096 per line, <= 4096 per tile) | records
            // two scans: (kept, records) packed and gc
            const uint64_t packed2 = ((uint64_t)ck << 40) | ((uint64_t)cg << 20) | (uint64_t)hdr_after;
            const uint64_t inc2 = warp_incl_scan_u64(packed2);
            const uint64_t ex2 = inc2 - packed2;
            const uint64_t tot2 = __shfl_sync(0xffffffffu, inc2, 31);
            (void)pk;
            const int64_t kept_before = run_kept + (int64_t)(ex2 >> 40);            // before this lane's line
            const int64_t gc_b = run_gc + (int64_t)((ex2 >> 20) & 0xFFFFF);
            const int64_t rec_before = run_rec + (int64_t)(ex2 & 0xFFFFF);          // records started before this lane's newline
            if (MODE == 1 && active) {
                if (line_is_hdr) {  // this newline ends the header of record rec_before - 1
                    const int64_t r = rec_before - 1;
                    if (r >= 0 && r < a.rec_cap) a.hdr_end[r] = tile_base + pos;
                }
                if (hdr_after) {  // a record starts right after this newline
                    const int64_t r = rec_before;
                    if (r < a.rec_cap) {
                        a.hdr_start[r] = tile_base + pos + 1;
                        a.seq_off[r] = kept_before + ck;
                        a.gc_prefix[r] = gc_b + cg;
                    } else {
                        a.result->overflow = 1;
                    }
                }
            }
            if (MODE == 2 && active && !line_is_hdr && len > 0) {  // copy the line to its place in the staging row
                const int dst = (int)(kept_before - in.kept_base) + (int)(in.kept_base & 15);
                const int src = ppos + 1;
                if (len <= 192u) fa_copy(s_out, sbytes, dst, src, (int)len, 0, 1);
            }
            if (MODE == 2) {  // long lines: the whole warp copies them, one after the other
                uint32_t longs = __ballot_sync(0xffffffffu, active && !line_is_hdr && len > 192u);
                while (longs) {
                    const int l = __ffs((int)longs) - 1;
                    longs &= longs - 1;
                    const int dst = (int)(__shfl_sync(0xffffffffu, kept_before, l) - in.kept_base) + (int)(in.kept_base & 15);
                    const int src = __shfl_sync(0xffffffffu, ppos, l) + 1;
                    const int ln = (int)__shfl_sync(0xffffffffu, len, l);
                    fa_copy(s_out, sbytes, dst, src, ln, lane, 32);
                }
            }
            run_kept += (int64_t)(tot2 >> 40);
            run_gc += (int64_t)((tot2 >> 20) & 0xFFFFF);
            run_rec += (int64_t)(tot2 & 0xFFFFF);
        }
        const int last = (n_events - lo > 32 ? 32 : n_events - lo) - 1;  // lane holding the last newline of this pass
        c_pos = __shfl_sync(0xffffffffu, pos, last);
        c_pg = __shfl_sync(0xffffffffu, pg, last);
        c_hdr = __shfl_sync(0xffffffffu, hdr_after, last);
    }

    // ---- C. the piece after the last newline
    int tail_len = g.data_end - (c_pos + 1);
    if (tail_len < 0) tail_len = 0;
    if (tail_len > 0 && sbytes[sidx(g.data_end - 1)] == '\r') {  // CR as the last valid byte: dropped iff a real LF follows in the input
        const int64_t nx = tile_base + g.data_end;
        if (nx < cr_limit && a.buf[nx] == '\n') tail_len -= 1;
    }
    const int tail_gc = total_g - c_pg;  // positions past data_end hold zeros (or the virtual newline): never G/C
    const bool has_nl = n_events > 0;

    if (MODE == 0) {
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) {
            f_nhs += __shfl_xor_sync(0xffffffffu, f_nhs, dd);
            f_kept += __shfl_xor_sync(0xffffffffu, f_kept, dd);
            f_gc += __shfl_xor_sync(0xffffffffu, f_gc, dd);
        }
        head_kept = __shfl_sync(0xffffffffu, head_kept, 0);
        head_gc = __shfl_sync(0xffffffffu, head_gc, 0);
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        if (has_nl) {
            if (!c_hdr) {
                f_kept += (uint32_t)tail_len;
                f_gc += (uint32_t)tail_gc;
            }
        } else {
            head_kept = (uint32_t)tail_len;
            head_gc = (uint32_t)tail_gc;
        }
        if (lane == 0) {
            FaTile t;
            const bool empty = g.data_end <= g.s0 && !g.virt;
            const bool ends_nl = has_nl && c_pos + 1 >= g.data_end;  // the last newline is the last valid position
            t.a = f_nhs | (has_nl ? FT_HAS_NL : 0u) | (c_hdr ? FT_HDR_LAST : 0u) | (ends_nl ? FT_ENDS_NL : 0u) | (b0_gt ? FT_B0_GT : 0u) |
                  (empty ? FT_EMPTY : 0u);
            t.rest = f_kept | (f_gc << 16);
            t.head = head_kept | (head_gc << 16);
            t.pos0 = pos0;
            *summary = t;
            if (!a.prev && tile == 0 && a.n > a.begin && !b0_gt) atomicMax(&a.result->err_pos, ~(unsigned long long)a.begin);  // data before the first '>'
        }
    }
    if (MODE == 2) {
        const bool tail_is_hdr = has_nl ? (c_hdr != 0) : (eff == ST_H);
        if (fast2) run_kept += WT_BYTES - n_events;  // everything but the LFs, already staged
        if (!fast2 && !tail_is_hdr && tail_len > 0) {
            const int dst = (int)(run_kept - in.kept_base) + (int)(in.kept_base & 15);
            const int src = c_pos + 1;
            fa_copy(s_out, sbytes, dst, src, tail_len, lane, 32);
            run_kept += tail_len;
        }
        __syncwarp();
        // flush the staging row: 16-byte chunks where they are whole, bytes at the two ragged ends
        const int shift = (int)(in.kept_base & 15);
        const int total = (int)(run_kept - in.kept_base);
        if (total > 0) {
            if (in.kept_base + total > a.seq_cap) {
                if (lane == 0) a.result->overflow = 1;
            } else {
                uint8_t* gdst = a.seq_out + (in.kept_base - shift);  // 16-byte aligned
                const int end = shift + total;
                for (int c = lane; c * 16 < end; c += 32) {
                    const int lo = c * 16, hi = lo + 16;
                    if (lo >= shift && hi <= end) {
                        *reinterpret_cast<uint4*>(gdst + lo) = *reinterpret_cast<const uint4*>(s_out + lo);
                    } else {
                        for (int i = (lo > shift ? lo : shift); i < (hi < end ? hi : end); i++) gdst[i] = s_out[i];
                    }
                }
            }
        }
    }
    __syncwarp();
}

// =================================================================== K1 / K3: TMA-fed persistent pass over all tiles
template <bool kCompact>
__global__ void __launch_bounds__(FA_THREADS) fasta_tile_kernel(const __grid_constant__ CUtensorMap tmap, const FastaScanArgs a, const uint32_t c7f,
                                                             const uint32_t c7b) {
    using SM = FaSmem<kCompact>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the warp index through a shuffle: the compiler then keeps tile arithmetic and the TMA operands in uniform registers
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u) != 0) __trap();
    uint8_t* data0 = smem_raw + warp * SM::per_warp;
    uint8_t* aux = smem_raw + SM::aux0 + warp * FaAux::total;
    uint8_t* s_out = kCompact ? smem_raw + ((SM::out0 + 15) & ~15) + warp * SM::out_bytes : nullptr;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(aux + FaAux::off_bar);
    const uint32_t data0_u32 = (uint32_t)__cvta_generic_to_shared(data0);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int64_t origin = a.begin & ~(int64_t)15;
    const int64_t n_tiles = a.n_tiles, full_rows = a.tma_rows;
    const int64_t stride = (int64_t)gridDim.x * FA_WARPS;
    auto issue = [&](int64_t tile, int b) {
        if (tile < n_tiles && tile * WT_ROWS < full_rows && lane == 0) {
            mbar_expect_tx(bar0 + 8 * b, WT_BYTES);
            tma_load_tile(data0_u32 + b * WT_BYTES, &tmap, (int)(tile * WT_ROWS), bar0 + 8 * b);
        }
    };
    int64_t cur = (int64_t)blockIdx.x * FA_WARPS + warp;
    issue(cur, 0);
    int b = 0;
    uint32_t phase_bits = 0;
    while (cur < n_tiles) {
        const int64_t nxt = cur + stride;
        issue(nxt, b ^ 1);
        uint8_t* sbytes = data0 + b * WT_BYTES;
        FaTileIn in{ST_S, 0, 0, 0, 0};
        FaTile k1_summary;
        if (kCompact) {  // what the scan left for this tile: fetched while the tile's bytes are still in flight
            const uint32_t ts = a.tile_state[cur];
            in.st = ts & 3u;
            in.regular = (ts >> 3) & 1u;
            in.kept_base = a.tile_base3[3 * cur + 1];
            k1_summary = a.tiles[cur];
        }
        if (cur * WT_ROWS < full_rows) {
            const uint32_t par = (phase_bits >> b) & 1u;
            while (!mbar_try_wait(bar0 + 8 * b, par)) {
            }
            phase_bits ^= 1u << b;
        }
        const FaGeom g = fa_geom(a, origin, cur);
        fa_fix_edges(a, sbytes, origin, cur, full_rows, lane, g);
        if (kCompact) {
            fa_tile<2>(a, sbytes, aux, origin, cur, g, c7f, c7b, in, &k1_summary, s_out);
        } else {
            fa_tile<0>(a, sbytes, aux, origin, cur, g, c7f, c7b, in, a.tiles + cur, nullptr);
        }
        cur = nxt;
        b ^= 1;
    }
}

// =================================================================== K2: per-record outputs of the tiles that have any
__global__ void __launch_bounds__(FA_THREADS) fasta_emit_kernel(const FastaScanArgs a, const uint32_t c7f, const uint32_t c7b) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* sbytes = smem_raw + warp * WT_BYTES;
    uint8_t* aux = smem_raw + FA_WARPS * WT_BYTES + warp * FaAux::total;
    const int64_t origin = a.begin & ~(int64_t)15;
    const int64_t n_tiles = a.n_tiles;
    const int64_t warps = (int64_t)gridDim.x * FA_WARPS;
    for (int64_t base = ((int64_t)blockIdx.x * FA_WARPS + warp) * 32; base < n_tiles; base += warps * 32) {
        const int64_t mine = base + lane;
        uint32_t todo = __ballot_sync(0xffffffffu, mine < n_tiles && (a.tile_state[mine] & 4u) != 0);
        while (todo) {
            const int64_t tile = base + (__ffs((int)todo) - 1);
            todo &= todo - 1;
            const int64_t tile_base = origin + tile * WT_BYTES;
            // stage the tile with plain loads into the same swizzled layout TMA produces
            for (int q = lane; q < WT_BYTES / 16; q += 32) {
                const int64_t p = tile_base + (int64_t)q * 16;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (p + 16 <= a.n) {
                    v = *reinterpret_cast<const uint4*>(a.buf + p);
                } else if (p < a.n) {
                    uint8_t tmp[16];
                    for (int i = 0; i < 16; i++) tmp[i] = p + i < a.n ? a.buf[p + i] : 0;
                    v = *reinterpret_cast<const uint4*>(tmp);
                }
                *reinterpret_cast<uint4*>(sbytes + sidx(q * 16)) = v;
            }
            __syncwarp();
            const FaGeom g = fa_geom(a, origin, tile);
            fa_fix_edges(a, sbytes, origin, tile, (int64_t)1 << 40, lane, g);  // everything is staged: no row is "missing"
            FaTileIn in;
            in.st = a.tile_state[tile] & 3u;
            in.rec_base = a.tile_base3[3 * tile + 0];
            in.kept_base = a.tile_base3[3 * tile + 1];
            in.gc_base = a.tile_base3[3 * tile + 2];
            fa_tile<1>(a, sbytes, aux, origin, tile, g, c7f, c7b, in, nullptr, nullptr);
            __syncwarp();
        }
    }
}

// =================================================================== the scan over tile summaries
constexpr int FS_THREADS = 256, FS_ITEMS = 8, FS_TILE = FS_THREADS * FS_ITEMS;

__device__ __forceinline__ FaAgg fs_thread_agg(const FaTile* tiles, int64_t first, int64_t n_tiles) {
    FaAgg acc = fa_identity();
    for (int i = 0; i < FS_ITEMS; i++) {
        const int64_t t = first + i;
        if (t < n_tiles) acc = fa_compose(acc, fa_of_tile(tiles[t]));
    }
    return acc;
}

// phase 1: one aggregate per block of FS_TILE tiles
__global__ void __launch_bounds__(FS_THREADS) fasta_scan_reduce_kernel(const FaTile* __restrict__ tiles, int64_t n_tiles, FaAgg* __restrict__ block_agg) {
    __shared__ FaAgg s_agg[FS_THREADS];
    const int t = threadIdx.x;
    s_agg[t] = fs_thread_agg(tiles, (int64_t)blockIdx.x * FS_TILE + (int64_t)t * FS_ITEMS, n_tiles);
    __syncthreads();
    if (t < 32) {
        FaAgg acc = fa_identity();
        for (int i = 0; i < FS_THREADS / 32; i++) acc = fa_compose(acc, s_agg[t * (FS_THREADS / 32) + i]);
        __syncwarp();
        s_agg[t] = acc;  // lanes own disjoint groups of 8 and group 0 starts at index 0: lane t's slot t was read by lane t / 8 <= t only
    }
    __syncthreads();
    if (t == 0) {
        FaAgg acc = fa_identity();
        for (int i = 0; i < 32; i++) acc = fa_compose(acc, s_agg[i]);
        block_agg[blockIdx.x] = acc;
    }
}

// phase 2: running value at the first tile of every block; totals into the result block.  One warp: the lanes fetch
// 32 block aggregates at a time, lane 0 applies them in order.
__global__ void __launch_bounds__(32) fasta_scan_spine_kernel(const FastaScanArgs a, const FaAgg* __restrict__ block_agg, int64_t n_blocks,
                                                              FaVal* __restrict__ block_val) {
    __shared__ FaAgg s_agg[32];
    __shared__ FaVal s_val[32];
    const int lane = threadIdx.x;
    FaVal v;
    v.st = ST_L;
    v.n = v.k = v.g = 0;
    if (lane == 0 && a.prev) {
        // chained range: the byte before `begin` says whether a line is open at all
        v.st = a.buf[a.begin - 1] == '\n' ? ST_L : (a.prev->tail_hdr ? ST_H : ST_S);
        v.n = a.prev->n_records;
        v.k = a.prev->seq_bytes;
        v.g = a.prev->gc_total;
        if (a.prev->err_pos != 0ull) atomicMax(&a.result->err_pos, a.prev->err_pos);
        if (a.prev->overflow) a.result->overflow = 1;
    }
    for (int64_t base = 0; base < n_blocks; base += 32) {
        if (base + lane < n_blocks) s_agg[lane] = block_agg[base + lane];
        __syncwarp();
        if (lane == 0) {
            const int m = n_blocks - base < 32 ? (int)(n_blocks - base) : 32;
            for (int i = 0; i < m; i++) {
                s_val[i] = v;
                v = fa_apply(v, s_agg[i]);
            }
        }
        __syncwarp();
        if (base + lane < n_blocks) block_val[base + lane] = s_val[lane];
        __syncwarp();
    }
    if (lane != 0) return;
    a.result->n_records = v.n;
    a.result->seq_bytes = v.k;
    a.result->gc_total = v.g;
    a.result->tail_hdr = v.st == ST_H ? 1ull : 0ull;
    if ((int64_t)v.n <= a.rec_cap) {
        a.seq_off[v.n] = (int64_t)v.k;
        a.gc_prefix[v.n] = (int64_t)v.g;
    } else {
        a.result->overflow = 1;
    }
}

// phase 3: state and bases at every tile's first byte; bit 2 of the state byte = the tile has per-record outputs
__global__ void __launch_bounds__(FS_THREADS) fasta_scan_down_kernel(const FaTile* __restrict__ tiles, int64_t n_tiles, const FaVal* __restrict__ block_val,
                                                                     uint8_t* __restrict__ tile_state, int64_t* __restrict__ tile_base3) {
    __shared__ FaAgg s_agg[FS_THREADS];
    __shared__ FaVal s_val[FS_THREADS];
    const int t = threadIdx.x;
    const int64_t first = (int64_t)blockIdx.x * FS_TILE + (int64_t)t * FS_ITEMS;
    s_agg[t] = fs_thread_agg(tiles, first, n_tiles);
    __syncthreads();
    if (t == 0) {  // 256 sequential applications: the data is tiny and this kernel is far off the critical path
        FaVal v = block_val[blockIdx.x];
        for (int i = 0; i < FS_THREADS; i++) {
            s_val[i] = v;
            v = fa_apply(v, s_agg[i]);
        }
    }
    __syncthreads();
    FaVal v = s_val[t];
    for (int i = 0; i < FS_ITEMS; i++) {
        const int64_t tl = first + i;
        if (tl >= n_tiles) break;
        const FaTile ft = tiles[tl];
        const uint32_t eff = ft_eff(ft.a, v.st);
        const bool emits = !(ft.a & FT_EMPTY) && (ft_nhs(ft.a, v.st) > 0 || (eff == ST_H && (ft.a & FT_HAS_NL)));
        tile_state[tl] = (uint8_t)(v.st | (emits ? 4u : 0u) | ((ft.a & FT_REGULAR) ? 8u : 0u));
        tile_base3[3 * tl + 0] = (int64_t)v.n;
        tile_base3[3 * tl + 1] = (int64_t)v.k;
        tile_base3[3 * tl + 2] = (int64_t)v.g;
        v = fa_apply(v, fa_of_tile(ft));
    }
}

// ------------------------------------------------------------------ launcher
int64_t fasta_scan_tiles(int64_t begin, int64_t n, int is_final) {
    const int64_t origin = begin & ~(int64_t)15;
    const int64_t t = (n + (is_final ? 1 : 0) - origin + WT_BYTES - 1) / WT_BYTES;
    return t > 0 ? t : 1;
}
int64_t fasta_workspace_payload(int64_t n_tiles) {
    const int64_t n_blocks = (n_tiles + FS_TILE - 1) / FS_TILE;
    return n_tiles * (int64_t)(sizeof(FaTile) + 1 + 24) + n_blocks * (int64_t)(sizeof(FaAgg) + sizeof(FaVal)) + 256;
}

template <bool kCompact>
static cudaError_t launch_fa_tile(FastaScanArgs a, cudaStream_t st) {
    constexpr int smem = FaSmem<kCompact>::cta_bytes;
    auto kern = fasta_tile_kernel<kCompact>;
    static int ctas_per_sm = 0, n_sm = 0;
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
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FA_THREADS, smem)) != cudaSuccess) return e;
        n_sm = sms;
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    alignas(64) CUtensorMap tm;
    if ((e = make_tensor_map(a.buf, a.begin, a.n, &tm, &a.tma_rows)) != cudaSuccess) return e;
    int64_t grid = (a.n_tiles + FA_WARPS - 1) / FA_WARPS;
    const int64_t persistent = (int64_t)n_sm * ctas_per_sm;
    if (grid > persistent) grid = persistent;
    kern<<<dim3((unsigned)grid), dim3(FA_THREADS), smem, st>>>(tm, a, 0x7F7F7F7Fu, 0x7B7B7B7Bu);
    return cudaGetLastError();
}

cudaError_t fasta_scan_launch(const FastaScanArgs& a0, int /*flags*/, cudaStream_t st) {
    FastaScanArgs a = a0;
    // workspace payload: FaTile[T] | base3 i64[3T] | FaAgg[B] | FaVal[B] | state u8[T]
    uint8_t* p = reinterpret_cast<uint8_t*>(a.payload);
    const int64_t T = a.n_tiles, B = (T + FS_TILE - 1) / FS_TILE;
    a.tiles = reinterpret_cast<FaTile*>(p);
    p += T * sizeof(FaTile);
    a.tile_base3 = reinterpret_cast<int64_t*>(p);
    p += T * 24;
    FaAgg* block_agg = reinterpret_cast<FaAgg*>(p);
    p += B * sizeof(FaAgg);
    FaVal* block_val = reinterpret_cast<FaVal*>(p);
    p += B * sizeof(FaVal);
    a.tile_state = p;
    cudaError_t e;
    if ((e = launch_fa_tile<false>(a, st)) != cudaSuccess) return e;
    fasta_scan_reduce_kernel<<<(unsigned)B, FS_THREADS, 0, st>>>(a.tiles, T, block_agg);
    fasta_scan_spine_kernel<<<1, 32, 0, st>>>(a, block_agg, B, block_val);
    fasta_scan_down_kernel<<<(unsigned)B, FS_THREADS, 0, st>>>(a.tiles, T, block_val, a.tile_state, a.tile_base3);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    {
        constexpr int smem = FA_WARPS * (WT_BYTES + FaAux::total);
        static std::atomic<unsigned> attr_done{0};  // per device, see launch_fa_tile
        int cur = 0;
        if ((e = cudaGetDevice(&cur)) != cudaSuccess) return e;
        const unsigned bit = 1u << (cur & 31);
        if (!(attr_done.load() & bit)) {
            if ((e = cudaFuncSetAttribute(fasta_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
            attr_done.fetch_or(bit);
        }
        int64_t blocks = (T + FA_WARPS * 32 - 1) / (FA_WARPS * 32);
        if (blocks > 148 * 8) blocks = 148 * 8;
        fasta_emit_kernel<<<(unsigned)blocks, FA_THREADS, smem, st>>>(a, 0x7F7F7F7Fu, 0x7B7B7B7Bu);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (a.seq_out) {
        if ((e = launch_fa_tile<true>(a, st)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace exb
