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
// kept sequence byte, G/C}; a decoupled look-back carries
//   {records so far, kept sequence bytes so far, G/C so far, "inside a header line"}
// so every record gets its sequence offset / G/C prefix as a plain store and the
// kept bytes are compacted straight into the output column at their final place
// (staged through shared memory so the global stores are 16-byte coalesced).
// gc_content per contig is then (gc_prefix[r+1]-gc_prefix[r]) / (seq_off[r+1]-seq_off[r])
// -- a 250 Mbp contig costs the same per byte as a 150 bp read.
#include "common.cuh"
#include "exon_b200_internal.h"

namespace exb {

struct alignas(16) FaState {
    uint64_t n_hdr;     // header lines started
    uint64_t seq;       // sequence bytes kept (aggregates: assuming the carried-in line is NOT a header)
    uint64_t gc;        // G/C among them (same assumption)
    uint64_t head_seq;  // aggregates: part of `seq` that lies before the first newline
    uint64_t head_gc;
    uint32_t has_nl;    // a newline was seen => tail_hdr is resolved (prefix states: always 1)
    uint32_t tail_hdr;  // the line open at the end is a header line
    uint64_t pad;
    __device__ static FaState combine(const FaState& p, const FaState& t) {
        FaState r;
        r.n_hdr = p.n_hdr + t.n_hdr;
        if (p.has_nl) {
            const bool h = p.tail_hdr != 0;
            r.seq = p.seq + t.seq - (h ? t.head_seq : 0);
            r.gc = p.gc + t.gc - (h ? t.head_gc : 0);
            r.head_seq = p.head_seq;
            r.head_gc = p.head_gc;
        } else {
            r.seq = p.seq + t.seq;
            r.gc = p.gc + t.gc;
            r.head_seq = p.head_seq + t.head_seq;
            r.head_gc = p.head_gc + t.head_gc;
        }
        r.has_nl = p.has_nl | t.has_nl;
        r.tail_hdr = t.has_nl ? t.tail_hdr : p.tail_hdr;
        r.pad = 0;
        return r;
    }
};
static_assert(sizeof(FaState) == 64, "FaState");

// five 12-bit fields in one scan word: header starts, kept, gc, head_kept, head_gc (each <= 2048 per warp)
__device__ __forceinline__ uint64_t pack5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
    return (uint64_t)a | ((uint64_t)b << 12) | ((uint64_t)c << 24) | ((uint64_t)d << 36) | ((uint64_t)e << 48);
}
struct U5 {
    int hs, kept, gc, hkept, hgc;
};
__device__ __forceinline__ U5 unpack5(uint64_t v) {
    U5 u;
    u.hs = (int)(v & 0xFFF);
    u.kept = (int)((v >> 12) & 0xFFF);
    u.gc = (int)((v >> 24) & 0xFFF);
    u.hkept = (int)((v >> 36) & 0xFFF);
    u.hgc = (int)((v >> 48) & 0xFFF);
    return u;
}

template <bool kCompact>
__global__ void __launch_bounds__(BLOCK_THREADS) fasta_scan_kernel(FastaScanArgs a) {
    __shared__ uint4 s_tile[TILE_CHUNKS];
    __shared__ uint8_t s_out[kCompact ? TILE_BYTES + 16 : 16];
    __shared__ uint64_t s_warp_tot[WARPS];
    __shared__ int s_warp_flag[WARPS];  // -1: no newline in the warp; else tail_hdr after its last newline
    __shared__ FaState s_excl;
    __shared__ int64_t s_tile_id;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint8_t* __restrict__ buf = a.buf;
    const int64_t origin = a.begin & ~(int64_t)15;

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

    // header starts: '>' right after a newline (or at `begin`)
    const bool fresh = a.prev == nullptr;  // `begin` is the start of the input (a line start)
    const int prev = (fresh && run_base == a.begin) ? '\n' : (run_base > a.begin || !fresh ? tile_byte(run_base - 1) : 0);
    const uint64_t after_nl = (pm << 1) | (prev == '\n' ? 1ull : 0ull);
    uint64_t hs = gt & after_nl & valid;
    if (fresh && a.begin >= run_base && a.begin < run_base + RUN_BYTES) {  // the byte at `begin` starts a line
        const int kb = (int)(a.begin - run_base);
        if ((gt >> kb) & 1ull) hs |= 1ull << kb;
        else if (a.n > a.begin) atomicMin(&a.result->err_pos, (unsigned long long)a.begin);  // data before the first '>'
    }
    // CR directly before a REAL LF is dropped (the virtual terminator at n strips nothing)
    const int next = tile_byte(run_base + RUN_BYTES);
    uint64_t real_nl = pm;
    if (a.is_final && a.n >= run_base && a.n < run_base + RUN_BYTES) real_nl &= ~(1ull << (int)(a.n - run_base));
    const bool next_is_real_nl = (next == '\n') && !(a.is_final && run_base + RUN_BYTES == a.n);
    const uint64_t crlf = crm & ((real_nl >> 1) | (next_is_real_nl ? (1ull << 63) : 0ull));

    // header spans that START in this run: from the '>' through the line's newline
    uint64_t hm = 0;
    int tail_hdr = 0;  // line open at the end of the run is a header (only meaningful if the run has a newline or a header start)
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
    const uint64_t head_span = low_bits64(first_nl == 64 ? 64 : first_nl + 1);  // bytes of the carried-in line (incl. its newline)
    // a header start inside the head region cannot exist (it needs a preceding newline) except at `begin`
    const uint64_t base_keep = ~pm & ~crlf & ~hm & valid;

    // ---- resolve "carried-in line is a header" from the nearest earlier newline / header start in the tile
    // a run changes the flag iff it contains a newline or a header start (hs at `begin` without newline)
    const bool sets_flag = cnt > 0 || hs != 0;
    const uint32_t setters = __ballot_sync(0xffffffffu, sets_flag);
    const uint32_t before = setters & ((1u << lane) - 1u);
    const int src = before ? 31 - __clz(before) : lane;
    const int src_flag = __shfl_sync(0xffffffffu, tail_hdr, src);
    if (lane == 0) s_warp_flag[warp] = -1;
    __syncwarp();
    if (setters && lane == 31 - __clz(setters)) s_warp_flag[warp] = tail_hdr;
    __syncthreads();
    int f_in = -1;  // -1 unresolved (depends on the tile's carry-in)
    if (before) f_in = src_flag;
    else
        for (int w = warp - 1; w >= 0; w--)
            if (s_warp_flag[w] >= 0) {
                f_in = s_warp_flag[w];
                break;
            }
    uint64_t keep = base_keep;
    uint32_t hkept = 0, hgc = 0;
    if (f_in == 1) keep &= ~head_span;
    else if (f_in < 0) {
        hkept = (uint32_t)__popcll(keep & head_span);
        hgc = (uint32_t)__popcll(keep & head_span & gm);
    }
    const uint32_t n_hs = (uint32_t)__popcll(hs);

    // ---- scans
    const uint64_t packed = pack5(n_hs, (uint32_t)__popcll(keep), (uint32_t)__popcll(keep & gm), hkept, hgc);
    const uint64_t incl = warp_incl_scan_u64(packed);
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    // cross-warp sums can exceed 12 bits, so unpack per warp and add as ints
    U5 ex = unpack5(incl - packed);
    {
        U5 o{0, 0, 0, 0, 0};
        for (int w = 0; w < warp; w++) {
            U5 u = unpack5(s_warp_tot[w]);
            o.hs += u.hs; o.kept += u.kept; o.gc += u.gc; o.hkept += u.hkept; o.hgc += u.hgc;
        }
        ex.hs += o.hs; ex.kept += o.kept; ex.gc += o.gc; ex.hkept += o.hkept; ex.hgc += o.hgc;
    }

    if (t == 0) {
        U5 tot{0, 0, 0, 0, 0};
        int last_flag = -1;
        for (int w = 0; w < WARPS; w++) {
            U5 u = unpack5(s_warp_tot[w]);
            tot.hs += u.hs; tot.kept += u.kept; tot.gc += u.gc; tot.hkept += u.hkept; tot.hgc += u.hgc;
            if (s_warp_flag[w] >= 0) last_flag = s_warp_flag[w];
        }
        FaState agg;
        agg.n_hdr = tot.hs;
        agg.seq = tot.kept;
        agg.gc = tot.gc;
        agg.head_seq = tot.hkept;
        agg.head_gc = tot.hgc;
        agg.has_nl = last_flag >= 0;
        agg.tail_hdr = last_flag > 0;
        agg.pad = 0;
        FaState init;
        init.n_hdr = 0; init.seq = 0; init.gc = 0; init.head_seq = 0; init.head_gc = 0;
        init.has_nl = 1; init.tail_hdr = 0; init.pad = 0;
        if (a.prev && tile == 0) {
            if (a.prev->err_pos != ~0ull) atomicMin(&a.result->err_pos, a.prev->err_pos);
            if (a.prev->overflow) a.result->overflow = 1;
            init.n_hdr = a.prev->n_records;
            init.seq = a.prev->seq_bytes;
            init.gc = a.prev->gc_total;
            init.tail_hdr = (uint32_t)a.prev->tail_hdr;
        }
        FaState excl = lookback<FaState>(a.slots, tile, agg, init);
        s_excl = excl;
        if (tile == a.n_tiles - 1) {
            FaState fin = FaState::combine(excl, agg);
            a.result->n_records = fin.n_hdr;
            a.result->seq_bytes = fin.seq;
            a.result->gc_total = fin.gc;
            a.result->tail_hdr = fin.tail_hdr;
            if ((int64_t)fin.n_hdr <= a.rec_cap) {
                a.seq_off[fin.n_hdr] = (int64_t)fin.seq;
                a.gc_prefix[fin.n_hdr] = (int64_t)fin.gc;
            } else {
                a.result->overflow = 1;
            }
        }
    }
    __syncthreads();

    // ---- finalise with the tile's carry-in
    const FaState excl = s_excl;
    const bool carry_hdr = excl.tail_hdr != 0;
    if (f_in < 0 && carry_hdr) keep &= ~head_span;
    const int64_t my_hdr = (int64_t)excl.n_hdr + ex.hs;
    const int64_t my_seq = (int64_t)excl.seq + ex.kept - (carry_hdr ? ex.hkept : 0);
    const int64_t my_gc = (int64_t)excl.gc + ex.gc - (carry_hdr ? ex.hgc : 0);
    const bool in_hdr_at_start = f_in < 0 ? carry_hdr : (f_in == 1);

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
        uint64_t hmask = hm | (in_hdr_at_start ? head_span : 0ull);
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
        // tile's kept count and output base after resolving the carry
        __shared__ int s_tile_kept;
        if (t == BLOCK_THREADS - 1) s_tile_kept = (ex.kept + __popcll(keep)) - (carry_hdr ? ex.hkept : 0);
        // NB: for the last thread ex.* are exclusive; its own head part is already removed from `keep`
        const int64_t obase = (int64_t)excl.seq;
        const int shift = (int)(obase & 15);
        int pos = shift + (int)(my_seq - obase);
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(s_tile);
        uint64_t m = keep;
        while (m) {
            const int k = __ffsll((long long)m) - 1;
            m &= m - 1;
            const int li = t * RUN_BYTES + k;
            s_out[pos++] = sb[swz(li >> 4) * 16 + (li & 15)];
        }
        __syncthreads();
        const int total = s_tile_kept;
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
