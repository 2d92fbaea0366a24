// inflate.cu -- BGZF (blocked gzip) input inflated ON THE DEVICE for sm_100a.
//
// Row SURVEY 8(f) rank 1 ("gzip / zstd input ... BGZF ... GPU inflate later").  The reference puts a streaming, single
// threaded GzipDecoder in front of the parser (rust/src/arrow_reader.rs:60-91 -> datafusion FileCompressionType::GZIP);
// bgzip'ed FASTQ / FASTA -- what sequencers and htslib tools write -- is a chain of INDEPENDENT gzip members of at most
// 64 KiB each, every one carrying its compressed size in the header and its uncompressed size + CRC-32 in the trailer
// (SAM specification 4.1).  So the host only walks the headers (exb_bgzf_index_host: 26 bytes looked at per block) and the
// compressed bytes cross PCIe as they are (1/3 - 1/4 of the text); one warp inflates one block (inflate_core.cuh) straight
// into the buffer the scan kernels read, then checks the block's CRC-32:
//
//   bgzf_inflate_kernel   grid = resident warps of the machine, blocks handed out by ticket (their cost varies);
//                         per warp 4 KB of shared memory (two Huffman index tables + canonical arrays)
//
// HBM traffic: compressed bytes once in, text once out (+ the match copies and the CRC pass, which hit L1 / L2: a block's
// 64 KiB window stays cached).  The kernel is bound by the serial symbol decode (instruction issue), not by DRAM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/exon_b200.h"
#include "inflate_core.cuh"

namespace exb {
int set_err(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

constexpr int IFL_WARPS = 8;  // warps per CTA: 8 x 4 KB of tables
#ifndef IFL_MINB
#define IFL_MINB 6            // resident CTAs per SM the kernel is compiled for (registers) and launched as (grid)
#endif
static_assert(sizeof(exb_bgzf_block) == 32, "exb_bgzf_block layout");

__global__ void __launch_bounds__(IFL_WARPS * 32, IFL_MINB) bgzf_inflate_kernel(const uint8_t* __restrict__ in, const exb_bgzf_block* __restrict__ blocks, int n_blocks,
                                                                       uint8_t* __restrict__ out, unsigned int* __restrict__ state, int check_crc, long long block_base) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ uint32_t s_crc[256];
    // lane and the warp's table pointer are computed ONCE: left to itself the compiler rebuilds both from %tid inside the
    // symbol loop (S2R + IMAD, 11 % of the instructions of the first version)
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    asm volatile("" : "+r"(lane));
    asm volatile("" : "+r"(warp));  // (the pointer itself stays a shared-memory pointer: LDS, not generic LD)
    ifl::Tables* tabs = reinterpret_cast<ifl::Tables*>(s_raw) + warp;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = ifl::crc_table_entry((uint32_t)i);
    __syncthreads();
    ifl::Tables& T = *tabs;
    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(&state[0], 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)n_blocks) break;
        const exb_bgzf_block blk = blocks[b];
        uint8_t* dst = out + blk.out_off;
        int err = ifl::inflate_block<32>(in + blk.in_off, (int)blk.clen, dst, (int)blk.isize, T, lane);
        if (err == ifl::OK && check_crc) {
            __syncwarp();  // the block's bytes were stored by all lanes
            const int n = (int)blk.isize;
            const int per = (n + 31) >> 5;
            const int lo = lane * per < n ? lane * per : n;
            const int mine = n - lo < per ? n - lo : per;
            const uint32_t c = ifl::crc_bytes(dst + lo, mine, s_crc);
            // combine in order (zlib crc32_combine): crc(A || B) = crc(A) * x^(8 |B|) + crc(B)
            const uint32_t x_per = ifl::crc_shift_bytes((uint32_t)per);
            uint32_t total = __shfl_sync(0xffffffffu, c, 0);
            for (int i = 1; i < 32; i++) {
                const uint32_t ci = __shfl_sync(0xffffffffu, c, i);
                const int ni = __shfl_sync(0xffffffffu, mine, i);
                if (ni == 0) break;
                total = ifl::crc_multmodp(ni == per ? x_per : ifl::crc_shift_bytes((uint32_t)ni), total) ^ ci;
            }
            if (total != blk.crc32) err = ifl::E_CRC;
        }
        if (err != ifl::OK && lane == 0) {
            // the FIRST failing block in file order is the one reported
            const unsigned long long w = ((unsigned long long)(block_base + b) << 8) | (unsigned)err;
            atomicMin(reinterpret_cast<unsigned long long*>(state + 2), w);
        }
    }
}

// state: [0] ticket (reset here), [2..3] min over failing members of (member << 8 | error), all ones = none: it is NOT reset
// here, so that a caller can let it accumulate over the launches of one file (block_base = members inflated before)
cudaError_t bgzf_inflate_launch(const uint8_t* d_in, const exb_bgzf_block* d_blocks, int64_t n_blocks, uint8_t* d_out, unsigned int* d_state, int check_crc,
                                int64_t block_base, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(d_state, 0, 4, st);
    if (e != cudaSuccess || n_blocks == 0) return e;
    const size_t dyn = IFL_WARPS * sizeof(ifl::Tables);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        e = cudaFuncSetAttribute(bgzf_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    int64_t ctas = (n_blocks + IFL_WARPS - 1) / IFL_WARPS;
    const int64_t machine = 148 * IFL_MINB;  // resident CTAs of 8 warps (33 KB of shared memory each)
    if (ctas > machine) ctas = machine;
    bgzf_inflate_kernel<<<(unsigned)ctas, IFL_WARPS * 32, dyn, st>>>(d_in, d_blocks, (int)n_blocks, d_out, d_state, check_crc, (long long)block_base);
    return cudaGetLastError();
}

}  // namespace exb

using namespace exb;

extern "C" {

// ---- host: walk the BGZF headers (no decompression)
static inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static inline uint32_t rd32(const uint8_t* p) { return rd16(p) | (rd16(p + 2) << 16); }

// the total size of the BGZF block whose header starts at p (n readable bytes), its payload offset; 0 = not a BGZF header
static int64_t bgzf_header(const uint8_t* p, int64_t n, int* payload_off) {
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const int64_t xlen = rd16(p + 10);
    if (12 + xlen > n) return 0;
    int64_t bsize = -1;
    for (int64_t o = 12; o + 4 <= 12 + xlen;) {  // subfields: SI1 SI2 SLEN data
        const int64_t slen = rd16(p + o + 2);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) bsize = (int64_t)rd16(p + o + 4) + 1;
        o += 4 + slen;
    }
    if (bsize < 0) return 0;
    // FNAME / FCOMMENT / FHCRC are not written by BGZF writers; a member that has them is left to the streaming decoder
    if (p[3] & ~4) return 0;
    *payload_off = (int)(12 + xlen);
    if (bsize < *payload_off + 8) return 0;
    return bsize;
}

int exb_bgzf_probe_host(const uint8_t* bytes, int64_t n) {
    int po = 0;
    return bytes && bgzf_header(bytes, n, &po) > 0 ? 1 : 0;
}

int exb_bgzf_index_host(const uint8_t* bytes, int64_t n, int64_t pos, int64_t max_out_bytes, exb_bgzf_block* blocks, int64_t max_blocks,
                        int64_t* n_blocks, int64_t* next_pos, int64_t* out_bytes) {
    if (!bytes || !blocks || !n_blocks || !next_pos || !out_bytes || pos < 0 || pos > n) return set_err(EXB_ERR_ARG, "exb_bgzf_index_host: bad arguments");
    int64_t k = 0, out = 0;
    const int64_t first = pos;
    while (pos < n && k < max_blocks) {
        int po = 0;
        const int64_t bs = bgzf_header(bytes + pos, n - pos, &po);
        if (bs == 0) return set_err(EXB_ERR_FORMAT, "not a BGZF block header at compressed offset %lld", (long long)pos);
        if (pos + bs > n) return set_err(EXB_ERR_FORMAT, "truncated BGZF block at compressed offset %lld", (long long)pos);
        const uint32_t isize = rd32(bytes + pos + bs - 4);
        if (isize > 65536u) return set_err(EXB_ERR_FORMAT, "BGZF block at compressed offset %lld claims %u bytes", (long long)pos, isize);
        if (k > 0 && out + isize > max_out_bytes) break;
        exb_bgzf_block& b = blocks[k++];
        b.in_off = pos - first + po;
        b.out_off = out;
        b.clen = (uint32_t)(bs - po - 8);
        b.isize = isize;
        b.crc32 = rd32(bytes + pos + bs - 8);
        b.reserved = 0;
        out += isize;
        pos += bs;
    }
    *n_blocks = k;
    *next_pos = pos;
    *out_bytes = out;
    return EXB_OK;
}

// ---- host: BGZF output (the writers' gzip sink).  Members of 65280 text bytes like htslib's, deflated by zlib on
// `threads` host threads; a member that does not fit 64 KiB (incompressible text) is stored.
static const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
constexpr int64_t BGZF_TEXT = 65280, BGZF_SLOT = 65536;

static int bgzf_member(const uint8_t* text, int n, int level, uint8_t* out) {  // returns the member's size, 0 on a zlib failure
    for (int attempt = 0; attempt < 2; attempt++) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, attempt ? 0 : level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return 0;
        zs.next_in = const_cast<Bytef*>(text);
        zs.avail_in = (uInt)n;
        zs.next_out = out + 18;
        zs.avail_out = (uInt)(BGZF_SLOT - 18 - 8);
        const int rc = deflate(&zs, Z_FINISH);
        const int clen = (int)zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) continue;  // did not fit: store it
        const int bsize = 18 + clen + 8;
        const uint8_t hdr[18] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, (uint8_t)((bsize - 1) & 0xFF), (uint8_t)((bsize - 1) >> 8)};
        memcpy(out, hdr, 18);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), text, (uInt)n);
        uint8_t* t = out + 18 + clen;
        for (int k = 0; k < 4; k++) t[k] = (uint8_t)(crc >> (8 * k)), t[4 + k] = (uint8_t)((uint32_t)n >> (8 * k));
        return bsize;
    }
    return 0;
}

int64_t exb_bgzf_compress_bound(int64_t n) { return ((n + BGZF_TEXT - 1) / BGZF_TEXT + 1) * BGZF_SLOT + 28; }

int exb_bgzf_compress_host(const uint8_t* text, int64_t n, int level, int threads, int eof_marker, uint8_t* out, int64_t out_cap, int64_t* out_bytes) {
    if (n < 0 || (n && !text) || !out || !out_bytes || out_cap < exb_bgzf_compress_bound(n)) return set_err(EXB_ERR_ARG, "exb_bgzf_compress_host: bad arguments");
    const int64_t pieces = (n + BGZF_TEXT - 1) / BGZF_TEXT;
    if (level < 0 || level > 9) level = 6;
    int T = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
    if (T > pieces) T = (int)(pieces > 0 ? pieces : 1);
    // every member is deflated into its own 64 KiB slot of `out` (the slots are then closed up in place)
    std::vector<int> size((size_t)pieces, 0);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        for (;;) {
            const int64_t k = next.fetch_add(1);
            if (k >= pieces) return;
            const int64_t lo = k * BGZF_TEXT;
            const int m = (int)(n - lo < BGZF_TEXT ? n - lo : BGZF_TEXT);
            size[(size_t)k] = bgzf_member(text + lo, m, level, out + k * BGZF_SLOT);
            if (size[(size_t)k] == 0) failed.store(1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
    if (failed.load()) return set_err(EXB_ERR_IO, "exb_bgzf_compress_host: deflate failed");
    int64_t w = 0;
    for (int64_t k = 0; k < pieces; k++) {
        if (w != k * BGZF_SLOT) memmove(out + w, out + k * BGZF_SLOT, (size_t)size[(size_t)k]);
        w += size[(size_t)k];
    }
    if (eof_marker) {
        memcpy(out + w, BGZF_EOF, sizeof(BGZF_EOF));
        w += sizeof(BGZF_EOF);
    }
    *out_bytes = w;
    return EXB_OK;
}

int64_t exb_bgzf_scratch_bytes(void) { return 16; }

int exb_bgzf_inflate(const uint8_t* d_in, const exb_bgzf_block* d_blocks, int64_t n_blocks, uint8_t* d_out, void* d_scratch, int check_crc,
                     void* stream_v) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    if (n_blocks < 0 || n_blocks > 0x7FFFFFFF || (n_blocks && (!d_in || !d_blocks || !d_out)) || !d_scratch)
        return set_err(EXB_ERR_ARG, "exb_bgzf_inflate: bad arguments");
    cudaError_t e = cudaMemsetAsync(d_scratch, 0xFF, 16, stream);  // no failing member yet
    if (e == cudaSuccess) e = bgzf_inflate_launch(d_in, d_blocks, n_blocks, d_out, reinterpret_cast<unsigned int*>(d_scratch), check_crc, 0, stream);
    return e == cudaSuccess ? EXB_OK : cuda_fail(e, "exb_bgzf_inflate");
}

static const char* ifl_err_text(int e) {
    switch (e) {
        case ifl::E_BTYPE: return "reserved block type";
        case ifl::E_STORED: return "stored block length mismatch";
        case ifl::E_HEADER: return "invalid dynamic Huffman header";
        case ifl::E_OVERSUB: return "over-subscribed Huffman code";
        case ifl::E_CODE: return "invalid code";
        case ifl::E_DIST: return "distance too far back";
        case ifl::E_OUTPUT: return "uncompressed size differs from the block trailer";
        case ifl::E_INPUT: return "compressed data ends early";
        case ifl::E_CRC: return "CRC-32 mismatch";
    }
    return "unknown error";
}

int exb_bgzf_status(const unsigned int* state_host, int64_t* bad_block) {
    const unsigned long long w = (unsigned long long)state_host[2] | ((unsigned long long)state_host[3] << 32);
    if (w == ~0ull) return EXB_OK;
    if (bad_block) *bad_block = (int64_t)(w >> 8);
    return set_err(EXB_ERR_FORMAT, "corrupt gzip stream: %s in BGZF block %lld", ifl_err_text((int)(w & 0xFF)), (long long)(w >> 8));
}

int exb_bgzf_finish(const void* d_scratch, int64_t* bad_block, void* stream_v) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    unsigned int st[4];
    cudaError_t e = cudaMemcpyAsync(st, d_scratch, sizeof(st), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "exb_bgzf_finish");
    return exb_bgzf_status(st, bad_block);
}

}  // extern "C"
