// common.cuh -- device helpers shared by the sm_100a scan kernels.
//
// Layout of one tile in the scan kernels
//   TILE_BYTES = 16 KiB of input = 1024 16-byte chunks.  256 threads; thread t
//   owns the CONTIGUOUS run of RUN_BYTES = 64 bytes [t*64, t*64+64) of the tile
//   so that one warp scan covers 2 KiB (the scans are amortised over 64 bytes
//   per lane instead of 16).  Global -> shared staging is done with coalesced
//   16-byte cp.async (LDGSTS) -- lane l of a warp copies chunk base+l -- and the
//   chunk index is XOR-swizzled on the way in so that the later per-thread
//   LDS.128 reads (stride 64 B between lanes) are bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exb {

constexpr int BLOCK_THREADS = 256;
constexpr int RUN_BYTES = 64;                            // bytes per thread
constexpr int RUN_CHUNKS = RUN_BYTES / 16;               // 16-byte chunks per thread
constexpr int TILE_BYTES = BLOCK_THREADS * RUN_BYTES;    // 16 KiB
constexpr int TILE_CHUNKS = TILE_BYTES / 16;             // 1024
constexpr int WARPS = BLOCK_THREADS / 32;

// ---------------------------------------------------------------- swizzle
// Logical chunk q -> physical chunk.  For thread t reading chunk j of its run,
// q = 4t + j; the 8 lanes of one LDS.128 phase then hit 8 distinct 16-byte
// bank groups.  A coalesced writer (32 consecutive q) still covers whole
// 128-byte rows, so the cp.async stores are conflict free too.
__device__ __forceinline__ int swz(int q) { return q ^ ((q >> 3) & 3); }

// ---------------------------------------------------------------- cp.async
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    // .cg: bypass L1 (streaming input, read once); bytes beyond src_bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Stage one tile: bytes [tile_base, tile_base + TILE_BYTES) of buf, clipped to
// [lo, hi).  Bytes outside the clip are zero (a neutral filler: not a newline,
// not G/C, adds 0 to byte sums).  buf must be 16-byte aligned.
__device__ __forceinline__ void stage_tile(uint4* s_tile, const uint8_t* __restrict__ buf, int64_t tile_base,
                                           int64_t lo, int64_t hi) {
#pragma unroll
    for (int i = 0; i < RUN_CHUNKS; i++) {
        int q = i * BLOCK_THREADS + threadIdx.x;
        int64_t g = tile_base + (int64_t)q * 16;
        int64_t rem = hi - g;
        int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        if (g < lo) nb = 0;  // partial leading chunks are patched by the caller (begin is chunk aligned in practice)
        const uint8_t* src = nb > 0 ? buf + g : buf;
        cp_async16(&s_tile[swz(q)], src, nb);
    }
    cp_async_commit();
}

// ---------------------------------------------------------------- byte tricks
// 0x80 in every byte of x that equals the byte replicated in pat (exact, no
// false positives).  3 ALU ops per word when pat/constants sit in registers.
__device__ __forceinline__ uint32_t eq_bytes(uint32_t x, uint32_t pat) {
    uint32_t t = (x & 0x7F7F7F7Fu) ^ pat;  // pat < 0x80 per byte
    return ~((t + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// 0x80 where byte is 'G' (0x47) or 'C' (0x43): they differ only in bit 2.
__device__ __forceinline__ uint32_t gc_bytes(uint32_t x) {
    uint32_t t = (x & 0x7B7B7B7Bu) ^ 0x43434343u;
    return ~((t + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// Gather the four 0x80 flags of m into a 4-bit mask (bit k = byte k).
// m*0x00204081 puts bit 7->28, 15->29, 23->30, 31->31 with no colliding terms.
__device__ __forceinline__ uint32_t nibble(uint32_t m) { return (m * 0x00204081u) >> 28; }

__device__ __forceinline__ uint32_t mask16(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    return nibble(m0) | (nibble(m1) << 4) | (nibble(m2) << 8) | (nibble(m3) << 12);
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4& v, uint32_t pat) {
    return mask16(eq_bytes(v.x, pat), eq_bytes(v.y, pat), eq_bytes(v.z, pat), eq_bytes(v.w, pat));
}
__device__ __forceinline__ uint32_t gc_mask16(const uint4& v) {
    return mask16(gc_bytes(v.x), gc_bytes(v.y), gc_bytes(v.z), gc_bytes(v.w));
}
// Sum of the 16 bytes of v as SIGNED chars (the reference's `c - 33` uses char).
__device__ __forceinline__ int sbyte_sum16(const uint4& v, int acc) {
    acc = __dp4a((int)v.x, 0x01010101, acc);
    acc = __dp4a((int)v.y, 0x01010101, acc);
    acc = __dp4a((int)v.z, 0x01010101, acc);
    acc = __dp4a((int)v.w, 0x01010101, acc);
    return acc;
}
// Signed-char sum of the first k (0..16) bytes of v.
__device__ __forceinline__ int sbyte_sum_prefix16(const uint4& v, int k) {
    // byte masks per word: word w keeps min(max(k-4w,0),4) low bytes
    auto keep = [](int c) -> uint32_t { return c <= 0 ? 0u : (c >= 4 ? 0xFFFFFFFFu : ((1u << (8 * c)) - 1u)); };
    int acc = 0;
    acc = __dp4a((int)(v.x & keep(k)), 0x01010101, acc);
    acc = __dp4a((int)(v.y & keep(k - 4)), 0x01010101, acc);
    acc = __dp4a((int)(v.z & keep(k - 8)), 0x01010101, acc);
    acc = __dp4a((int)(v.w & keep(k - 12)), 0x01010101, acc);
    return acc;
}

__device__ __forceinline__ uint64_t low_bits64(int k) {  // k in 0..64 -> mask of k low bits
    return k >= 64 ? ~0ull : ((1ull << k) - 1ull);
}

// ---------------------------------------------------------------- memory ordering
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg_u4(uint4* p, const uint4& v) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------- decoupled look-back
// Single-pass chained scan over tiles (Merrill & Garland).  State is a POD of
// 32 bytes with an associative combine(prefix, next).  Each tile publishes its
// own aggregate (flag 1) as soon as it is known, then walks back over its
// predecessors until it meets an inclusive prefix (flag 2), publishes its own
// inclusive prefix and continues.  Tile ids come from an atomic ticket so a
// tile's predecessors are always resident or finished (forward progress).
constexpr uint32_t FLAG_EMPTY = 0, FLAG_AGG = 1, FLAG_INC = 2;

struct alignas(16) TileSlot {  // 144 bytes: states of up to 64 bytes
    uint4 agg[4];
    uint4 inc[4];
    uint32_t flag;
    uint32_t pad[3];
};

template <typename State>
__device__ __forceinline__ void slot_store(uint4* dst, const State& s) {
    static_assert(sizeof(State) % 16 == 0 && sizeof(State) <= 64, "State must be 16..64 bytes");
    const uint4* p = reinterpret_cast<const uint4*>(&s);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(State) / 16); i++) st_cg_u4(dst + i, p[i]);
}
template <typename State>
__device__ __forceinline__ State slot_load(const uint4* src) {
    State s;
    uint4* p = reinterpret_cast<uint4*>(&s);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(State) / 16); i++) p[i] = ld_cg_u4(src + i);
    return s;
}

// Called by ONE thread.  `agg` is this tile's aggregate, `init` the state before
// tile 0.  Returns the exclusive prefix (state before this tile).
template <typename State>
__device__ __forceinline__ State lookback(TileSlot* slots, int64_t tile, const State& agg, const State& init) {
    if (tile == 0) {
        State inc = State::combine(init, agg);
        slot_store(slots[0].inc, inc);
        __threadfence();
        st_release_u32(&slots[0].flag, FLAG_INC);
        return init;
    }
    slot_store(slots[tile].agg, agg);
    __threadfence();
    st_release_u32(&slots[tile].flag, FLAG_AGG);

    State suffix;  // aggregate of tiles (k, tile)
    bool have = false;
    int64_t k = tile - 1;
    State excl;
    while (true) {
        uint32_t f;
        do {
            f = ld_acquire_u32(&slots[k].flag);
        } while (f == FLAG_EMPTY);
        State s = slot_load<State>(f == FLAG_INC ? slots[k].inc : slots[k].agg);
        suffix = have ? State::combine(s, suffix) : s;
        have = true;
        if (f == FLAG_INC) {
            excl = suffix;
            break;
        }
        k--;
    }
    State inc = State::combine(excl, agg);
    slot_store(slots[tile].inc, inc);
    __threadfence();
    st_release_u32(&slots[tile].flag, FLAG_INC);
    return excl;
}

// ---------------------------------------------------------------- block-wide look-back over plain sums
// The chained state is NW independent 64-bit words per tile, each carrying its own
// 2-bit status in the top bits (value < 2^62), so one load both polls and fetches
// (no flag-then-payload round trip) and a word can never be seen half-written.
// All BLOCK_THREADS threads take part: thread i inspects predecessor tile-1-i, so a
// round covers 256 predecessors in one L2 round trip.  The prefix frontier can then
// advance by 256 tiles per round trip instead of 1 (serial walk) or 32 (one warp):
// at 16 KiB per tile that is the difference between 0.4 TB/s and not being the limit.
constexpr uint64_t CH_AGG = 1ull << 62, CH_INC = 2ull << 62, CH_VAL = (1ull << 62) - 1ull;

__device__ __forceinline__ uint64_t ld_acquire_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

template <int NW>
struct LookbackSmem {
    uint64_t sum[NW][WARPS];      // sum of the aggregates nearer than the warp's first inclusive word
    uint64_t inc_val[NW][WARPS];  // that inclusive word's value
    uint32_t ready[NW][WARPS];    // no needed predecessor of this warp is still unpublished
    uint32_t has_inc[NW][WARPS];
};

// Every thread of the block calls this with the same arguments; every thread gets the
// exclusive prefix (sum over tiles [0, tile) plus init).  Thread 0 publishes.  Anything
// the block wrote to global memory before the call is visible to whoever later observes
// this tile's words (release), and everything published by observed tiles is visible
// after the call (acquire + barrier).
template <int NW>
__device__ __forceinline__ void block_lookback(uint64_t* chain, int64_t tile, const uint64_t (&agg)[NW], const uint64_t (&init)[NW],
                                               uint64_t (&excl)[NW], LookbackSmem<NW>* sm) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (tile == 0) {
        if (t == 0) {
            __threadfence();
#pragma unroll
            for (int j = 0; j < NW; j++) st_release_u64(&chain[j], CH_INC | ((init[j] + agg[j]) & CH_VAL));
        }
#pragma unroll
        for (int j = 0; j < NW; j++) excl[j] = init[j];
        return;
    }
    if (t == 0) {
        __threadfence();
#pragma unroll
        for (int j = 0; j < NW; j++) st_release_u64(&chain[tile * NW + j], CH_AGG | (agg[j] & CH_VAL));
    }
    uint64_t acc[NW];
    bool done[NW];
#pragma unroll
    for (int j = 0; j < NW; j++) {
        acc[j] = 0;
        done[j] = false;
    }
    int64_t base = tile - 1;
    while (true) {
        const int64_t idx = base - t;
#pragma unroll
        for (int j = 0; j < NW; j++) {
            if (done[j]) continue;
            uint64_t w;
            if (idx >= 0) w = ld_acquire_u64(&chain[idx * NW + j]);
            else w = CH_INC | (idx == -1 ? (init[j] & CH_VAL) : 0ull);  // the state before tile 0 acts as an inclusive word
            const uint32_t flag = (uint32_t)(w >> 62);
            const uint64_t val = w & CH_VAL;
            const uint32_t inc_mask = __ballot_sync(0xffffffffu, flag == 2);
            const uint32_t empty_mask = __ballot_sync(0xffffffffu, flag == 0);
            const int first = inc_mask ? __ffs(inc_mask) - 1 : 32;
            const uint32_t needed = first >= 32 ? 0xffffffffu : ((1u << first) - 1u);
            uint64_t part = lane < first ? val : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            const uint64_t iv = __shfl_sync(0xffffffffu, val, first & 31);
            if (lane == 0) {
                sm->sum[j][warp] = part;
                sm->inc_val[j][warp] = iv;
                sm->ready[j][warp] = (empty_mask & needed) == 0;
                sm->has_inc[j][warp] = inc_mask != 0;
            }
        }
        __syncthreads();
        bool all_ready = true, all_done = true;
        uint64_t add[NW];
        bool fin[NW];
#pragma unroll
        for (int j = 0; j < NW; j++) {
            add[j] = 0;
            fin[j] = false;
            if (done[j]) continue;
            bool ready = true;
            for (int w = 0; w < WARPS; w++) {
                if (!sm->ready[j][w]) {
                    ready = false;
                    break;
                }
                add[j] += sm->sum[j][w];
                if (sm->has_inc[j][w]) {
                    add[j] += sm->inc_val[j][w];
                    fin[j] = true;
                    break;
                }
            }
            if (!ready) all_ready = false;
        }
#pragma unroll
        for (int j = 0; j < NW; j++) {
            if (done[j]) continue;
            if (fin[j]) {  // found its inclusive word inside this window: final
                acc[j] += add[j];
                done[j] = true;
            } else if (all_ready) {
                acc[j] += add[j];  // whole window were aggregates: keep walking
                all_done = false;
            } else {
                all_done = false;
            }
        }
        __syncthreads();  // sm is rewritten by the next round
        if (all_done) break;
        if (all_ready) base -= BLOCK_THREADS;  // else: poll the same window again
    }
#pragma unroll
    for (int j = 0; j < NW; j++) excl[j] = acc[j];
    if (t == 0) {
#pragma unroll
        for (int j = 0; j < NW; j++) st_release_u64(&chain[tile * NW + j], CH_INC | ((acc[j] + agg[j]) & CH_VAL));
    }
}

// One-word look-back run by ONE warp (all 32 lanes call it; the other warps of the block
// wait at a barrier and execute nothing meanwhile).  Each lane inspects two predecessors,
// so a round covers 64 tiles.  The caller has already published this tile's aggregate
// (CH_AGG | agg) -- as early as it was known -- and tile 0 publishes CH_INC itself.
// Aggregates must fit 32 bits (they are per-tile counts).  Returns the exclusive prefix
// in every lane and publishes the inclusive word.
__device__ __forceinline__ uint64_t warp_lookback(uint64_t* chain, int64_t tile, uint64_t agg, uint64_t init) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) return init;
    uint64_t acc = 0;
    int64_t base = tile - 1;
    while (true) {
        const int64_t i0 = base - lane, i1 = base - 32 - lane;
        const uint64_t w0 = i0 >= 0 ? ld_acquire_u64(&chain[i0]) : (CH_INC | (i0 == -1 ? (init & CH_VAL) : 0ull));
        const uint64_t w1 = i1 >= 0 ? ld_acquire_u64(&chain[i1]) : (CH_INC | (i1 == -1 ? (init & CH_VAL) : 0ull));
        const uint32_t f0 = (uint32_t)(w0 >> 62), f1 = (uint32_t)(w1 >> 62);
        const uint32_t inc0 = __ballot_sync(0xffffffffu, f0 == 2), emp0 = __ballot_sync(0xffffffffu, f0 == 0);
        const int first0 = inc0 ? __ffs(inc0) - 1 : 32;
        const uint32_t need0 = first0 >= 32 ? 0xffffffffu : ((1u << first0) - 1u);
        if (emp0 & need0) continue;  // a predecessor nearer than the first inclusive word has not published yet: poll again
        acc += __reduce_add_sync(0xffffffffu, lane < first0 ? (uint32_t)w0 : 0u);
        if (inc0) {
            acc += __shfl_sync(0xffffffffu, w0 & CH_VAL, first0);
            break;
        }
        const uint32_t inc1 = __ballot_sync(0xffffffffu, f1 == 2), emp1 = __ballot_sync(0xffffffffu, f1 == 0);
        const int first1 = inc1 ? __ffs(inc1) - 1 : 32;
        const uint32_t need1 = first1 >= 32 ? 0xffffffffu : ((1u << first1) - 1u);
        if (emp1 & need1) {  // first window was all aggregates and is consumed; re-poll from the second one
            base -= 32;
            continue;
        }
        acc += __reduce_add_sync(0xffffffffu, lane < first1 ? (uint32_t)w1 : 0u);
        if (inc1) {
            acc += __shfl_sync(0xffffffffu, w1 & CH_VAL, first1);
            break;
        }
        base -= 64;
    }
    if (lane == 0) st_release_u64(&chain[tile], CH_INC | ((acc + agg) & CH_VAL));
    return acc;
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

}  // namespace exb
