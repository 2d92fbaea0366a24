// peer_exchange.cu -- the shard-edge exchange of a byte-range sharded scan over NVLink peer memory (sm_100a).
//
// What the ranks of one box must agree on is tiny: one 128-byte scan result block per shard (newline count,
// open line, tail sums: the input of the exact FASTQ phase resync) and 8 int64 aggregates (COUNT / sums).  Through NCCL
// that is an all-gather + an all-reduce, two collectives of ~100-200 us each at 8 ranks -- a quarter of a 1.4 ms scan
// step.  Here every rank owns a small SYMMETRIC buffer (same layout on every GPU, mapped into every peer through
// NVLink / NVSwitch); a rank stores its block straight into each peer's buffer, raises a per-sender sequence flag with a
// system-scope release, and spins on its own flags.  One launch of one CTA per exchange, ~10 us.
//
// Layout of a rank's buffer (EXB_PEER_BYTES):
//   [   0,  128)  gather flags  u64[16]   flag[j] = last step for which rank j's block has landed here
//   [ 128,  256)  reduce flags  u64[16]
//   [ 256, 4352)  blocks[2][16][128 B]    double-buffered by step parity: a rank one step ahead never overwrites
//   [4352, 6400)  aggs  [2][16][ 8 x i64] what a slower peer is still reading (it cannot be two steps ahead: it
//                                         would need this rank's flag of the step in between)
//   [6400, 6408)  status u64              1 = a wait timed out (a peer died); sticky
//   [6528, 6656)  record flags u64[16]    (single-exchange COUNT, see peer_count_fused_kernel)
//   [8192,16384)  records[2][16][256 B]
// Steps are numbered from 1 on the host (seq); the buffer is zeroed and the ranks barrier once before step 1.
#include <cuda_runtime.h>
#include <stdint.h>

#include "exon_b200_internal.h"
#include "combine_records.cuh"

namespace exb {

constexpr int PX_MAX_WORLD = 16;
constexpr int PX_GFLAG = 0, PX_RFLAG = 128, PX_BLOCKS = 256, PX_AGGS = PX_BLOCKS + 2 * PX_MAX_WORLD * 128;
constexpr int PX_STATUS = PX_AGGS + 2 * PX_MAX_WORLD * 64;
constexpr int PX_FFLAG = 6528, PX_RECS = 8192;
constexpr int PX_BYTES = 16384;
constexpr unsigned long long PX_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}
// lane 0 of warp j waits until rank j's flag in MY buffer has reached seq
__device__ __forceinline__ bool wait_flag(const uint64_t* flag, uint64_t seq) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys_u64(flag) < seq) {
        if (globaltimer_ns() - t0 > PX_TIMEOUT_NS) return false;
        __nanosleep(100);
    }
    return true;
}

// warp j: my block -> rank j's blocks[parity][rank]; flag; then wait for rank j's block in my own buffer.
__global__ void __launch_bounds__(32 * PX_MAX_WORLD) peer_allgather_kernel(uint8_t* const* __restrict__ peers, int rank, int world,
                                                                         const uint8_t* __restrict__ block, uint64_t seq) {
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int parity = (int)(seq & 1);
    uint8_t* mine = peers[rank];
    if (j < world) {
        uint8_t* dst = peers[j] + PX_BLOCKS + (parity * PX_MAX_WORLD + rank) * 128;
        if (lane < 8) reinterpret_cast<uint4*>(dst)[lane] = reinterpret_cast<const uint4*>(block)[lane];
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            st_release_sys_u64(reinterpret_cast<uint64_t*>(peers[j] + PX_GFLAG) + rank, seq);
            if (!wait_flag(reinterpret_cast<const uint64_t*>(mine + PX_GFLAG) + j, seq)) *reinterpret_cast<volatile uint64_t*>(mine + PX_STATUS) = 1;
        }
    }
}

// This shard's contribution to the global aggregates of a COUNT query -- [0..5] the scan's aggregates, [6] lines of the
// file mod 4 (reported by the last shard; must be 0), [7] 1 if this shard met a malformed record -- stored into every
// peer's aggs[parity][rank]; then the sum over ranks in rank order (int64: order-independent, but fixed anyway).
__global__ void __launch_bounds__(32 * PX_MAX_WORLD) peer_count_reduce_kernel(uint8_t* const* __restrict__ peers, int rank, int world,
                                                                            const uint64_t* __restrict__ result_block,
                                                                            const int64_t* __restrict__ agg, int is_last, uint64_t seq,
                                                                            int64_t* __restrict__ total) {
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int parity = (int)(seq & 1);
    uint8_t* mine = peers[rank];
    if (j < world) {
        if (lane < 8) {
            int64_t v;
            if (lane < 6) v = agg[lane];
            else if (lane == 6) v = is_last ? (int64_t)(result_block[0] & 3ull) : 0;  // exb_scan_result.total_lines
            else v = result_block[2] != 0 ? 1 : 0;  // err_pos word as stored on the device: 0 = no malformed record
            reinterpret_cast<int64_t*>(peers[j] + PX_AGGS + (parity * PX_MAX_WORLD + rank) * 64)[lane] = v;
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            st_release_sys_u64(reinterpret_cast<uint64_t*>(peers[j] + PX_RFLAG) + rank, seq);
            if (!wait_flag(reinterpret_cast<const uint64_t*>(mine + PX_RFLAG) + j, seq)) *reinterpret_cast<volatile uint64_t*>(mine + PX_STATUS) = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int64_t* a = reinterpret_cast<const int64_t*>(mine + PX_AGGS + parity * PX_MAX_WORLD * 64);
        int64_t s = 0;
        for (int r = 0; r < world; r++) s += *reinterpret_cast<const volatile int64_t*>(a + r * 8 + threadIdx.x);
        if (threadIdx.x == 7 && *reinterpret_cast<volatile uint64_t*>(mine + PX_STATUS) != 0) s = -1;  // exchange failed: poison the check word
        total[threadIdx.x] = s;
    }
}

// The sharded COUNT with ONE exchange (fastq_scan.cu, fastq_fused_candidates_kernel): warp j stores this rank's 256-byte
// record into rank j's records[parity][rank], raises the flag, waits for rank j's record in its own buffer; then one
// thread combines the `world` records -- every rank computes the same global aggregates, so nothing has to be reduced.
// Replaces peer_allgather_kernel -> fastq_compose_prev_kernel -> fastq_fused_combine_kernel -> peer_count_reduce_kernel.
__global__ void __launch_bounds__(32 * PX_MAX_WORLD) peer_count_fused_kernel(uint8_t* const* __restrict__ peers, int rank, int world,
                                                                           const uint8_t* __restrict__ record, const int64_t* __restrict__ ranges,
                                                                           FusedPreds fp, uint64_t seq, long long* __restrict__ total) {
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int parity = (int)(seq & 1);
    uint8_t* mine = peers[rank];
    if (j < world) {
        uint8_t* dst = peers[j] + PX_RECS + (parity * PX_MAX_WORLD + rank) * 256;
        if (lane < 16) reinterpret_cast<uint4*>(dst)[lane] = reinterpret_cast<const uint4*>(record)[lane];
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            st_release_sys_u64(reinterpret_cast<uint64_t*>(peers[j] + PX_FFLAG) + rank, seq);
            if (!wait_flag(reinterpret_cast<const uint64_t*>(mine + PX_FFLAG) + j, seq)) *reinterpret_cast<volatile uint64_t*>(mine + PX_STATUS) = 1;
        }
    }
    __syncthreads();
    // the records were written by peers over NVLink: every thread fetches a few words past L1 (volatile) after the acquire on
    // their flags (one thread reading all 32 x world words one after the other cost ~40 us), then one thread combines
    __shared__ long long s_recs[PX_MAX_WORLD * 32];
    __shared__ int64_t s_ranges[PX_MAX_WORLD * 3];
    {
        const volatile long long* src = reinterpret_cast<const volatile long long*>(mine + PX_RECS + parity * PX_MAX_WORLD * 256);
        for (int i = threadIdx.x; i < world * 32; i += blockDim.x) s_recs[i] = src[i];
        for (int i = threadIdx.x; i < world * 3; i += blockDim.x) s_ranges[i] = ranges[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fastq_combine_records(s_recs, s_ranges, world, fp.p, fp.n, total);
        if (*reinterpret_cast<volatile uint64_t*>(mine + PX_STATUS) != 0) total[7] = -1;  // exchange failed: poison the check word
    }
}

int set_err(int code, const char* fmt, ...);

}  // namespace exb

using namespace exb;

extern "C" {

int64_t exb_peer_bytes(void) { return PX_BYTES; }
int64_t exb_peer_blocks_offset(uint64_t seq) { return PX_BLOCKS + (int64_t)(seq & 1) * PX_MAX_WORLD * 128; }

int exb_peer_allgather_block(void* const* d_peers, int rank, int world, const void* d_block, uint64_t seq, void* stream) {
    if (world < 1 || world > PX_MAX_WORLD || rank < 0 || rank >= world || seq == 0)
        return set_err(EXB_ERR_ARG, "exb_peer_allgather_block: world must be 1..%d, 0 <= rank < world, seq >= 1", PX_MAX_WORLD);
    peer_allgather_kernel<<<1, 32 * world, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint8_t* const*>(d_peers), rank, world,
                                                                     reinterpret_cast<const uint8_t*>(d_block), seq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(EXB_ERR_CUDA, "peer_allgather launch: %s", cudaGetErrorString(e));
    return 0;
}

int exb_peer_count_reduce(void* const* d_peers, int rank, int world, const void* d_workspace, const int64_t* d_agg, int is_last, uint64_t seq,
                          int64_t* d_total, void* stream) {
    if (world < 1 || world > PX_MAX_WORLD || rank < 0 || rank >= world || seq == 0)
        return set_err(EXB_ERR_ARG, "exb_peer_count_reduce: world must be 1..%d, 0 <= rank < world, seq >= 1", PX_MAX_WORLD);
    peer_count_reduce_kernel<<<1, 32 * world, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint8_t* const*>(d_peers), rank, world,
                                                                        reinterpret_cast<const uint64_t*>(d_workspace), d_agg, is_last, seq, d_total);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(EXB_ERR_CUDA, "peer_count_reduce launch: %s", cudaGetErrorString(e));
    return 0;
}

int exb_peer_count_fused(void* const* d_peers, int rank, int world, const void* d_record, const int64_t* d_ranges, const exb_predicate* preds,
                         int n_preds, uint64_t seq, int64_t* d_total, void* stream) {
    if (world < 1 || world > PX_MAX_WORLD || rank < 0 || rank >= world || seq == 0 || n_preds < 0 || n_preds > EXB_MAX_PREDICATES || !d_record ||
        !d_ranges || !d_total)
        return set_err(EXB_ERR_ARG, "exb_peer_count_fused: bad arguments (world 1..%d, seq >= 1)", PX_MAX_WORLD);
    FusedPreds fp;
    fp.n = n_preds;
    for (int i = 0; i < n_preds; i++) fp.p[i] = preds[i];
    peer_count_fused_kernel<<<1, 32 * world, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint8_t* const*>(d_peers), rank, world,
                                                                        reinterpret_cast<const uint8_t*>(d_record), d_ranges, fp, seq,
                                                                        reinterpret_cast<long long*>(d_total));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(EXB_ERR_CUDA, "peer_count_fused launch: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
