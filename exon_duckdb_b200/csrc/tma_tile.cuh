// tma_tile.cuh -- the warp-tile skeleton shared by the FASTQ and FASTA byte passes (sm_100a).
//
// The input is viewed as rows of 128 bytes (a 2-D uint8 tensor map over the file image, no copy).  A warp owns a
// 4 KiB tile = 32 rows; ONE elected lane issues ONE cp.async.bulk.tensor.2d (UTMALDG) per tile with the 128-byte
// swizzle, completing on an mbarrier; lane l then owns row l and reads it with eight conflict-free LDS.128.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace exb {

constexpr int WT_BYTES = 4096;  // bytes per warp tile = 32 rows of 128 bytes, one row per lane
constexpr int WT_ROWS = 32;
constexpr int ROW_BYTES = 128;

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
// 32-bit equality mask of two consecutive 16-byte chunks.  Each IDP.4A chain leaves (8 flag bits) << 7; two multiply-adds
// and one shift put the four pieces in place (the 16-bit form above would cost two more instructions per 32 bytes).
__device__ __forceinline__ uint32_t flags_to_x16(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {  // (16-bit mask) << 7
    uint32_t lo = __dp4a(m0, 0x08040201u, 0u);
    lo = __dp4a(m1, 0x80402010u, lo);
    uint32_t hi = __dp4a(m2, 0x08040201u, 0u);
    hi = __dp4a(m3, 0x80402010u, hi);
    return hi * 256u + lo;
}
__device__ __forceinline__ uint32_t nl_mask32r(const uint4& a, const uint4& b, uint32_t c7f, uint32_t pat) {
    const uint32_t x0 = flags_to_x16(eq_flags(a.x, c7f, pat), eq_flags(a.y, c7f, pat), eq_flags(a.z, c7f, pat), eq_flags(a.w, c7f, pat));
    const uint32_t x1 = flags_to_x16(eq_flags(b.x, c7f, pat), eq_flags(b.y, c7f, pat), eq_flags(b.z, c7f, pat), eq_flags(b.w, c7f, pat));
    return x1 * 512u + (x0 >> 7);
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

__device__ __forceinline__ uint32_t gc_mask32r(const uint4& a, const uint4& b, uint32_t c7b, uint32_t c7f, uint32_t pat) {
    const uint32_t x0 = flags_to_x16(gc_flags(a.x, c7b, c7f, pat), gc_flags(a.y, c7b, c7f, pat), gc_flags(a.z, c7b, c7f, pat), gc_flags(a.w, c7b, c7f, pat));
    const uint32_t x1 = flags_to_x16(gc_flags(b.x, c7b, c7f, pat), gc_flags(b.y, c7b, c7f, pat), gc_flags(b.z, c7b, c7f, pat), gc_flags(b.w, c7b, c7f, pat));
    return x1 * 512u + (x0 >> 7);
}

// ---------------------------------------------------------------- TMA / mbarrier
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// one 32 x 128 B box of the input (rows [row0, row0 + 32)), swizzled 128B, completing `bytes` on the barrier
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* tmap, int row0, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
                 "l"(tmap), "r"(mbar), "r"(0), "r"(row0)
                 : "memory");
}

// tile-local byte index -> byte offset in the 128B-swizzled tile buffer
__device__ __forceinline__ int sidx(int li) {
    return (li & 0xF8F) | ((((li >> 4) ^ (li >> 7)) & 7) << 4);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// The input viewed as rows of 128 bytes starting at `origin`; only rows that lie completely inside the range.
static inline cudaError_t make_tensor_map(const uint8_t* buf, int64_t begin, int64_t n, CUtensorMap* tm, int64_t* rows_out) {
    const int64_t origin = begin & ~(int64_t)15;
    int64_t rows = (n - origin) / ROW_BYTES;
    if (rows < 0) rows = 0;
    *rows_out = rows;
    memset(tm, 0, sizeof(*tm));
    if (rows == 0) return cudaSuccess;  // tiny input: the kernel stages it by hand
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)ROW_BYTES, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ROW_BYTES};
    const cuuint32_t box[2] = {(cuuint32_t)ROW_BYTES, (cuuint32_t)WT_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(buf) + origin, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}


}  // namespace exb
