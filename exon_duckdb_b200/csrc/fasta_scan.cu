// fasta_scan.cu -- single-pass FASTA record scan + sequence compaction for sm_100a.
//
// Replaces noodles-fasta 0.27 Reader::read_definition / read_sequence as driven
// by exon 0.2.6's FASTA batch reader (reference call sites: rust/src/
// arrow_reader.rs:104-118,125-153).  Semantics:
//   * a record starts at a line whose FIRST byte is '>' ('>' elsewhere is data);
//   * the sequence is every byte of the following non-header lines, with each
//     line's LF (and a CR directly before that LF) removed; blank lines add nothing.
// Every input byte is read once.  Per tile (16 KiB in shared memory) each thread
// classifies its 64-byte run into bit masks {newline, header start, header span,
// kept sequence byte, G/C}.  Two things cross tile boundaries:
//   (1) "is the line that is open at my first byte a header line?" -- local to the
//       nearest predecessor that contains a line start, so every tile publishes a
//       one-word record {sets the flag, flag} with no dependency and reads its
//       predecessor's (walking further back only across tiles without a newline);
//   (2) three plain running sums {records, kept sequence bytes, G/C among them},
//       chained with the block-wide decoupled look-back.
// So every record gets its sequence offset / G/C prefix as a plain store and the
// kept bytes are compacted straight into the output column at their final place
// (staged through shared memory so the global stores are 16-byte coalesced).
// gc_content per contig is then (gc_prefix[r+1]-gc_prefix[r]) / (seq_off[r+1]-seq_off[r])
// -- a 250 Mbp contig costs the same per byte as a 150 bp read.
#include "common.cuh"
#include "exon_b200_internal.h"

namespace exb {

// three 20-bit fields in one scan word: header starts, kept, gc (each <= 2048 per warp)
__device__ __forceinline__ uint64_t pack3u(uint32_t a, uint32_t b, uint32_t c) {
    return (uint64_t)a | ((uint64_t)b << 20) | ((uint64_t)c << 40);
}
struct U3 {
    int hs, kept, gc;
};
__device__ __forceinline__ U3 unpack3u(uint64_t v) {
    U3 u;
    u.hs = (int)(v & 0xFFFFF);
    u.kept = (int)((v >> 20) & 0xFFFFF);
    u.gc = (int)((v >> 40) & 0xFFFFF);
    return u;
}

constexpr uint64_t REC_VALID = 1ull << 63;

template <bool kCompact>
__global__ void __launch_bounds__(BLOCK_THREADS) fasta_scan_kernel(FastaScanArgs a) {
    __shared__ uint4 s_tile[TILE_CHUNKS];
    __shared__ __align__(16) uint8_t s_out[kCompact ? TILE_BYTES + 16 : 16];
    __shared__ uint64_t s_warp_tot[WARPS];
    __shared__ int s_warp_flag[WARPS];  // -1: the warp's 2 KiB has no line start; else "open line is a header" at its end
    __shared__ LookbackSmem<3> s_lb;
    __shared__ int s_carry;
    __shared__ int64_t s_tile_id;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint8_t* __restrict__ buf = a.buf;
    const int64_t origin = a.begin & ~(int64_t)15;
    uint64_t* chain = reinterpret_cast<uint64_t*>(a.slots);
    uint64_t* recs = chain + 3 * a.n_tiles;

    if (t == 0) s_tile_id = (int64_t)atomicAdd(a.ticket, 1ull);
    __syncthreads();
    const int64_t tile = s_tile_id;
    const int64_t tile_base = origin + tile * TILE_BYTES;

    stage_tile(s_tile, buf, tile_base, origin, a.n);
    cp_async_wait<0>();
    __syncthreads();

    const bool has_begin_pad = (tile == 0 && a.begin != origin);
    const bool has_eof = a.is_final && (a.n >= tile_base && a.n < tile_base + TILE_BYTES);
    if (has_begin_pad || has_eof) {
        if (t == 0) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(s_tile);
            if (has_begin_pad)
                for (int64_t i = origin; i < a.begin; i++) sb[swz(0) * 16 + (int)(i - origin)] = 0;
            if (has_eof && a.n > a.begin && buf[a.n - 1] != '\n') {  // virtual terminator for the last line
                int li = (int)(a.n - tile_base);
                sb[swz(li >> 4) * 16 + (li & 15)] = '\n';
            }
        }
        __syncthreads();
    }
    auto tile_byte = [&](int64_t abs_pos) -> int {
        int64_t li = abs_pos - tile_base;
        if (li >= 0 && li < TILE_BYTES && (abs_pos < a.n || (a.is_final && abs_pos == a.n))) {
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(s_tile);
            return sb[swz((int)(li >> 4)) * 16 + (int)(li & 15)];
        }
        return (abs_pos >= (a.prev ? 0 : a.begin) && abs_pos < a.halo_n) ? (int)buf[abs_pos] : -1;
    };

    // ---- masks over the thread's 64-byte run
    const int64_t run_base = tile_base + (int64_t)t * RUN_BYTES;
    uint64_t pm, gt, crm, gm;
    {
        uint4 c0 = s_tile[swz(4 * t + 0)], c1 = s_tile[swz(4 * t + 1)];
        uint4 c2 = s_tile[swz(4 * t + 2)], c3 = s_tile[swz(4 * t + 3)];
        auto m64 = [&](uint32_t pat) -> uint64_t {
            uint32_t lo = eq_mask16(c0, pat) | (eq_mask16(c1, pat) << 16);
            uint32_t hi = eq_mask16(c2, pat) | (eq_mask16(c3, pat) << 16);
            return ((uint64_t)hi << 32) | lo;
        };
        pm = m64(0x0A0A0A0Au);
        gt = m64(0x3E3E3E3Eu);
        crm = m64(0x0D0D0D0Du);
        uint32_t glo = gc_mask16(c0) | (gc_mask16(c1) << 16);
        uint32_t ghi = gc_mask16(c2) | (gc_mask16(c3) << 16);
        gm = ((uint64_t)ghi << 32) | glo;
    }
    // bytes of [begin, n): filler outside must never count as sequence
    uint64_t valid = ~0ull;
    if (run_base + RUN_BYTES > a.n) valid = run_base >= a.n ? 0ull : low_bits64((int)(a.n - run_base));
    if (run_base < a.begin) valid &= (a.begin - run_base >= 64) ? 0ull : ~low_bits64((int)(a.begin - run_base));

    // header starts: '>' right after a newline (or at `begin` of an unchained range)
    const bool fresh = a.prev == nullptr;
    const int prev = (fresh && run_base == a.begin) ? '\n' : ((run_base > a.begin || !fresh) ? tile_byte(run_base - 1) : 0);
    const uint64_t after_nl = (pm << 1) | (prev == '\n' ? 1ull : 0ull);
    uint64_t hs = gt & after_nl & valid;
    if (fresh && a.begin >= run_base && a.begin < run_base + RUN_BYTES) {  // the byte at `begin` starts a line
        const int kb = (int)(a.begin - run_base);
        if ((gt >> kb) & 1ull) hs |= 1ull << kb;
        else if (a.n > a.begin) atomicMax(&a.result->err_pos, ~(unsigned long long)a.begin);  // data before the first '>'
    }
    // CR directly before a REAL LF is dropped (the virtual terminator at n strips nothing)
    const int next = tile_byte(run_base + RUN_BYTES);
    uint64_t real_nl = pm;
    if (a.is_final && a.n >= run_base && a.n < run_base + RUN_BYTES) real_nl &= ~(1ull << (int)(a.n - run_base));
    const bool next_is_real_nl = (next == '\n') && !(a.is_final && run_base + RUN_BYTES == a.n);
    const uint64_t crlf = crm & ((real_nl >> 1) | (next_is_real_nl ? (1ull << 63) : 0ull));

    // header spans that START in this run: from the '>' through the line's newline
    uint64_t hm = 0;
    int tail_hdr = 0;  // the line open at the end of the run is a header
    {
        uint64_t s = hs;
        while (s) {
            const int h = __ffsll((long long)s) - 1;
            s &= s - 1;
            const uint64_t above = pm & ~low_bits64(h);
            if (above) {
                const int e = __ffsll((long long)above) - 1;
                hm |= low_bits64(e + 1) & ~low_bits64(h);
            } else {
                hm |= ~low_bits64(h);
                tail_hdr = 1;
            }
        }
    }
    const int cnt = __popcll(pm);
    const int first_nl = cnt ? __ffsll((long long)pm) - 1 : 64;
    const uint64_t head_span = low_bits64(first_nl == 64 ? 64 : first_nl + 1);  // the carried-in line (incl. its newline)
    uint64_t keep = ~pm & ~crlf & ~hm & valid;

    // ---- "carried-in line is a header": nearest earlier run that contains a line start decides
    const bool sets_flag = cnt > 0 || hs != 0;
    const uint32_t setters = __ballot_sync(0xffffffffu, sets_flag);
    const uint32_t before = setters & ((1u << lane) - 1u);
    const int src = before ? 31 - __clz(before) : lane;
    const int src_flag = __shfl_sync(0xffffffffu, tail_hdr, src);
    if (lane == 0) s_warp_flag[warp] = -1;
    __syncwarp();
    if (setters && lane == 31 - __clz(setters)) s_warp_flag[warp] = tail_hdr;
    __syncthreads();
    if (t == 0) {
        // publish this tile's record first (it depends on nobody), then resolve the carry from the predecessors'
        int last = -1;
        for (int w = 0; w < WARPS; w++)
            if (s_warp_flag[w] >= 0) last = s_warp_flag[w];
        __threadfence();
        st_release_u64(&recs[tile], REC_VALID | (last >= 0 ? 2ull : 0ull) | (last > 0 ? 1ull : 0ull));
        int carry = -1;
        for (int64_t k = tile - 1; k >= 0 && carry < 0; k--) {
            uint64_t w;
            do {
                w = ld_acquire_u64(&recs[k]);
            } while (!(w & REC_VALID));
            if (w & 2ull) carry = (int)(w & 1ull);
        }
        if (carry < 0) carry = a.prev ? (int)a.prev->tail_hdr : 0;
        s_carry = carry;
        if (a.prev && tile == 0) {
            if (a.prev->err_pos != 0ull) atomicMax(&a.result->err_pos, a.prev->err_pos);
            if (a.prev->overflow) a.result->overflow = 1;
        }
    }
    __syncthreads();
    int f_in;
    if (before) f_in = src_flag;
    else {
        f_in = s_carry;
        for (int w = warp - 1; w >= 0; w--)
            if (s_warp_flag[w] >= 0) {
                f_in = s_warp_flag[w];
                break;
            }
    }
    if (f_in) keep &= ~head_span;
    const uint32_t n_hs = (uint32_t)__popcll(hs);

    // ---- scans of (records, kept bytes, G/C)
    const uint64_t packed = pack3u(n_hs, (uint32_t)__popcll(keep), (uint32_t)__popcll(keep & gm));
    const uint64_t incl = warp_incl_scan_u64(packed);
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint64_t wsum = 0, tsum = 0;  // 20-bit fields: eight warps of <= 2048 each cannot overflow
    for (int w = 0; w < WARPS; w++) {
        if (w < warp) wsum += s_warp_tot[w];
        tsum += s_warp_tot[w];
    }
    const U3 ex = unpack3u(wsum + incl - packed);
    const U3 tot = unpack3u(tsum);
    const uint64_t agg[3] = {(uint64_t)tot.hs, (uint64_t)tot.kept, (uint64_t)tot.gc};
    const uint64_t init[3] = {a.prev ? a.prev->n_records : 0ull, a.prev ? a.prev->seq_bytes : 0ull, a.prev ? a.prev->gc_total : 0ull};
    uint64_t excl[3];
    block_lookback<3>(chain, tile, agg, init, excl, &s_lb);

    if (tile == a.n_tiles - 1 && t == 0) {
        const uint64_t n_hdr = excl[0] + tot.hs, seq = excl[1] + tot.kept, gc = excl[2] + tot.gc;
        int last = s_carry;
        for (int w = 0; w < WARPS; w++)
            if (s_warp_flag[w] >= 0) last = s_warp_flag[w];
        a.result->n_records = n_hdr;
        a.result->seq_bytes = seq;
        a.result->gc_total = gc;
        a.result->tail_hdr = (uint64_t)last;
        if ((int64_t)n_hdr <= a.rec_cap) {
            a.seq_off[n_hdr] = (int64_t)seq;
            a.gc_prefix[n_hdr] = (int64_t)gc;
        } else {
            a.result->overflow = 1;
        }
    }

    const int64_t my_hdr = (int64_t)excl[0] + ex.hs;
    const int64_t my_seq = (int64_t)excl[1] + ex.kept;
    const int64_t my_gc = (int64_t)excl[2] + ex.gc;

    // record starts
    {
        uint64_t s = hs;
        int rank = 0;
        while (s) {
            const int k = __ffsll((long long)s) - 1;
            s &= s - 1;
            const int64_t r = my_hdr + rank++;
            if (r < a.rec_cap) {
                a.hdr_start[r] = run_base + k;
                a.seq_off[r] = my_seq + __popcll(keep & low_bits64(k));
                a.gc_prefix[r] = my_gc + __popcll(keep & gm & low_bits64(k));
            } else {
                a.result->overflow = 1;
            }
        }
    }
    // header line terminators
    {
        uint64_t hmask = hm | (f_in ? head_span : 0ull);
        uint64_t s = pm & hmask;
        while (s) {
            const int k = __ffsll((long long)s) - 1;
            s &= s - 1;
            const int64_t r = my_hdr + __popcll(hs & low_bits64(k)) - 1;
            if (r >= 0 && r < a.rec_cap) a.hdr_end[r] = run_base + k;
        }
    }

    // ---- compaction of the kept bytes
    if (kCompact) {
        const int64_t obase = (int64_t)excl[1];
        const int shift = (int)(obase & 15);
        int pos = shift + ex.kept;
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(s_tile);
        uint64_t m = keep;
        while (m) {
            const int k = __ffsll((long long)m) - 1;
            m &= m - 1;
            const int li = t * RUN_BYTES + k;
            s_out[pos++] = sb[swz(li >> 4) * 16 + (li & 15)];
        }
        __syncthreads();
        const int total = tot.kept;
        if (obase + total > a.seq_cap) {
            if (t == 0) a.result->overflow = 1;
        } else {
            uint8_t* gdst = a.seq_out + (obase - shift);  // 16-byte aligned
            const int end = shift + total;
            const int nchunks = (end + 15) >> 4;
            for (int c = t; c < nchunks; c += BLOCK_THREADS) {
                const int lo = c * 16, hi = lo + 16;
                if (lo >= shift && hi <= end) {
                    *reinterpret_cast<uint4*>(gdst + lo) = *reinterpret_cast<const uint4*>(s_out + lo);
                } else {
                    for (int i = max(lo, shift); i < min(hi, end); i++) gdst[i] = s_out[i];
                }
            }
        }
    }
}

cudaError_t fasta_scan_launch(const FastaScanArgs& a, int /*flags*/, cudaStream_t st) {
    dim3 grid((unsigned)a.n_tiles), block(BLOCK_THREADS);
    if (a.seq_out)
        fasta_scan_kernel<true><<<grid, block, 0, st>>>(a);
    else
        fasta_scan_kernel<false><<<grid, block, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace exb
