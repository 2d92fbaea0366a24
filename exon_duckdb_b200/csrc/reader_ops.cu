// reader_ops.cu -- the last kernels of a chunk in the native reader (reader.cu): they turn the per-record arrays and the
// gathered columns into exactly what the host's vectors hold, so that the DuckDB scan callback assigns pointers instead
// of looping over rows.
//   string_t_kernel      Arrow-style offsets + bytes -> DuckDB string_t entries (duckdb string_type.hpp:19-60); replaces
//                        the SetVectorString loop of the reference's consumer (duckdb arrow_conversion.cpp:252-266)
//   string_t_borrow_kernel  the same for columns that are byte ranges of the input chunk: the entries point into the pinned
//                        INPUT block on the host, no bytes are gathered or copied back
//   valid_bits_kernel    one validity byte per row -> DuckDB / Arrow validity bitmap (arrow_conversion.cpp:36-76)
//   list_entries_kernel  list offsets -> list_entry_t {offset, length} relative to each 2048-row batch (types.hpp:56-68)
//   computed columns     gc_content (sequence_functions/module.cpp:131-158), list_avg(quality_score_string_to_list(q))
//                        (fastq_functions/module.cpp:32-50 + DuckDB's avg finalise, see x87div.h), byte lengths
#include "common.cuh"
#include "exon_b200_internal.h"
#include "x87div.h"

namespace exb {

static int grid_for(int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    if (b > 148 * 32) b = 148 * 32;
    if (b < 1) b = 1;
    return (int)b;
}

// off[i] index `data` (device copy of the column, same layout as the host buffer the D2H copy fills);
// host_base = host address of data[0].
__global__ void __launch_bounds__(256) string_t_kernel(const int64_t* __restrict__ off, const uint8_t* __restrict__ data, uint64_t host_base,
                                                       int64_t n, uint4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = off[i];
        const uint32_t len = (uint32_t)(off[i + 1] - b);
        const uint8_t* p = data + b;
        uint32_t w[3] = {0u, 0u, 0u};
        const uint32_t take = len <= 12u ? len : 4u;  // inlined value, or the 4-byte prefix
        for (uint32_t k = 0; k < take; k++) w[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
        uint4 v;
        v.x = len;
        v.y = w[0];
        if (len <= 12u) {
            v.z = w[1];
            v.w = w[2];
        } else {
            const uint64_t ptr = host_base + (uint64_t)b;
            v.z = (uint32_t)ptr;
            v.w = (uint32_t)(ptr >> 32);
        }
        out[i] = v;
    }
}
cudaError_t string_t_launch(const int64_t* off, const uint8_t* data, uint64_t host_base, int64_t n, void* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    string_t_kernel<<<grid_for(n, 256), 256, 0, st>>>(off, data, host_base, n, reinterpret_cast<uint4*>(out));
    return cudaGetLastError();
}

// The same entries for a column whose rows are byte ranges of the INPUT chunk (every FASTQ column, FASTA id / description):
// start[i] / len[i] index the device copy of the chunk, host_base = host address of the chunk's byte 0 in the pinned
// block the IO thread filled.  Strings above 12 bytes then point INTO THE INPUT BLOCK on the host: the column's bytes
// are neither gathered on the device nor copied back -- only these 16 bytes per row cross PCIe.
__global__ void __launch_bounds__(256) string_t_borrow_kernel(const int64_t* __restrict__ start, const uint32_t* __restrict__ len_of,
                                                              const uint8_t* __restrict__ data, uint64_t host_base, int64_t n, uint4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = start[i];
        const uint32_t len = len_of[i];
        const uint8_t* p = data + b;
        uint32_t w[3] = {0u, 0u, 0u};
        const uint32_t take = len <= 12u ? len : 4u;  // inlined value, or the 4-byte prefix
        for (uint32_t k = 0; k < take; k++) w[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
        uint4 v;
        v.x = len;
        v.y = w[0];
        if (len <= 12u) {
            v.z = w[1];
            v.w = w[2];
        } else {
            const uint64_t ptr = host_base + (uint64_t)b;
            v.z = (uint32_t)ptr;
            v.w = (uint32_t)(ptr >> 32);
        }
        out[i] = v;
    }
}
cudaError_t string_t_borrow_launch(const int64_t* start, const uint32_t* len, const uint8_t* data, uint64_t host_base, int64_t n, void* out,
                                   cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    string_t_borrow_kernel<<<grid_for(n, 256), 256, 0, st>>>(start, len, data, host_base, n, reinterpret_cast<uint4*>(out));
    return cudaGetLastError();
}

// words = ceil(n / 64); rows past n read as valid (DuckDB ignores them, Arrow requires nothing of them)
__global__ void __launch_bounds__(256) valid_bits_kernel(const uint8_t* __restrict__ valid, int64_t n, unsigned long long* __restrict__ out,
                                                         unsigned long long* __restrict__ null_count) {
    const int64_t words = (n + 63) >> 6;
    unsigned long long nulls = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long m = 0;
        const int64_t base = w << 6;
        for (int k = 0; k < 64; k++) {
            const int64_t i = base + k;
            const bool v = i < n ? valid[i] != 0 : true;
            m |= (unsigned long long)v << k;
        }
        out[w] = m;
        nulls += 64 - __popcll(m);
    }
    if (null_count && nulls) atomicAdd(null_count, nulls);
}
cudaError_t valid_bits_launch(const uint8_t* valid, int64_t n, uint64_t* out, uint64_t* null_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    valid_bits_kernel<<<grid_for((n + 63) / 64, 256), 256, 0, st>>>(valid, n, reinterpret_cast<unsigned long long*>(out),
                                                                     reinterpret_cast<unsigned long long*>(null_count));
    return cudaGetLastError();
}

// entries[i] = {off[i] - off[first row of i's batch], off[i+1] - off[i]};  bases[k] = off[k * batch_rows] - off[0] (first
// child value of batch k), closed with the total
__global__ void __launch_bounds__(256) list_entries_kernel(const int64_t* __restrict__ off, int64_t n, int64_t batch_rows,
                                                           ulonglong2* __restrict__ out, long long* __restrict__ bases) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = off[(i / batch_rows) * batch_rows];
        out[i] = make_ulonglong2((unsigned long long)(off[i] - base), (unsigned long long)(off[i + 1] - off[i]));
        if (i % batch_rows == 0) bases[i / batch_rows] = base - off[0];
        if (i == n - 1) bases[(n + batch_rows - 1) / batch_rows] = off[n] - off[0];
    }
}
cudaError_t list_entries_launch(const int64_t* off, int64_t n, int64_t batch_rows, void* out, int64_t* bases, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    list_entries_kernel<<<grid_for(n, 256), 256, 0, st>>>(off, n, batch_rows, reinterpret_cast<ulonglong2*>(out), reinterpret_cast<long long*>(bases));
    return cudaGetLastError();
}

// gc_content per selected record: FASTQ from the scan's per-record counts, FASTA from its prefix arrays
__global__ void __launch_bounds__(256) gc_sel_kernel(const uint32_t* __restrict__ seq_len, const uint32_t* __restrict__ gc,
                                                     const int64_t* __restrict__ seq_off, const int64_t* __restrict__ gc_prefix,
                                                     const int64_t* __restrict__ sel, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = sel ? sel[i] : i;
        long long len, g;
        if (seq_len) {
            len = seq_len[r];
            g = gc[r];
        } else {
            len = seq_off[r + 1] - seq_off[r];
            g = gc_prefix[r + 1] - gc_prefix[r];
        }
        // (float)gc_count / (float)size with C int -> float conversions and one IEEE divide (module.cpp:148-156)
        out[i] = len == 0 ? 0.0f : __fdiv_rn(__ll2float_rn(g), __ll2float_rn(len));
    }
}
cudaError_t gc_sel_launch(const uint32_t* seq_len, const uint32_t* gc, const int64_t* seq_off, const int64_t* gc_prefix, const int64_t* sel,
                          int64_t n, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    gc_sel_kernel<<<grid_for(n, 256), 256, 0, st>>>(seq_len, gc, seq_off, gc_prefix, sel, n, out);
    return cudaGetLastError();
}

// list_avg(quality_score_string_to_list(q)): (double)((long double)sum / (long double)n), NULL for an empty list
__global__ void __launch_bounds__(256) mean_quality_kernel(const uint32_t* __restrict__ qual_len, const int32_t* __restrict__ qsum,
                                                           const int64_t* __restrict__ sel, int64_t n, double* __restrict__ out,
                                                           uint8_t* __restrict__ valid) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = sel ? sel[i] : i;
        const uint32_t len = qual_len[r];
        out[i] = len ? exb_x87_div((int64_t)qsum[r], len) : 0.0;
        valid[i] = len ? 1 : 0;
    }
}
cudaError_t mean_quality_launch(const uint32_t* qual_len, const int32_t* qsum, const int64_t* sel, int64_t n, double* out, uint8_t* valid,
                                cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    mean_quality_kernel<<<grid_for(n, 256), 256, 0, st>>>(qual_len, qsum, sel, n, out, valid);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) lens_i64_kernel(const uint32_t* __restrict__ lens, int64_t n, long long* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (long long)lens[i];
}
cudaError_t lens_i64_launch(const uint32_t* lens, int64_t n, int64_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    lens_i64_kernel<<<grid_for(n, 256), 256, 0, st>>>(lens, n, reinterpret_cast<long long*>(out));
    return cudaGetLastError();
}

// What the host needs to know about a scanned FASTQ chunk before it can go on, in ONE small copy:
//   info[0] = offset just past the last complete record (0 if there is none)
//   info[1] = number of complete records
__global__ void fastq_chunk_info_kernel(const ScanResult* __restrict__ res, const uint32_t* __restrict__ line_end, long long* __restrict__ info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long R = res->total_lines / 4;
    info[1] = (long long)R;
    info[0] = (R > 0 && !res->overflow) ? (long long)line_end[4 * R - 1] + 1 : 0;
}
cudaError_t fastq_chunk_info_launch(const void* ws, const uint32_t* line_end, int64_t* info, cudaStream_t st) {
    fastq_chunk_info_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const ScanResult*>(ws), line_end, reinterpret_cast<long long*>(info));
    return cudaGetLastError();
}

// COUNT / SUM / AVG over the per-record arrays (BASELINE C5): sum len, sum gc, sum round(gc_content * 2^32)
__global__ void __launch_bounds__(256) seq_totals_kernel(const uint32_t* __restrict__ seq_len, const uint32_t* __restrict__ gc, int64_t n,
                                                         unsigned long long* __restrict__ totals) {
    unsigned long long sl = 0, sg = 0, sf = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t l = seq_len[i], g = gc[i];
        sl += l;
        sg += g;
        if (l) sf += (unsigned long long)__double2ll_rn((double)__fdiv_rn(__uint2float_rn(g), __uint2float_rn(l)) * 4294967296.0);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sl += __shfl_xor_sync(0xffffffffu, sl, d);
        sg += __shfl_xor_sync(0xffffffffu, sg, d);
        sf += __shfl_xor_sync(0xffffffffu, sf, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sl) atomicAdd(totals + 1, sl);
        if (sg) atomicAdd(totals + 2, sg);
        if (sf) atomicAdd(totals + 5, sf);
    }
}
cudaError_t seq_totals_launch(const uint32_t* seq_len, const uint32_t* gc, int64_t n, int64_t* totals, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    seq_totals_kernel<<<grid_for(n, 1024), 256, 0, st>>>(seq_len, gc, n, reinterpret_cast<unsigned long long*>(totals));
    return cudaGetLastError();
}

// Small read-backs (scan result block, column totals, counts) go to MAPPED pinned host memory through a kernel instead of
// cudaMemcpyAsync: a 16-byte device-to-host copy queues behind the 64 MiB result copy of the previous chunk in the same
// copy engine and waited ~1.7 ms per chunk for it (profiles/round2_reader_trace.txt).  dst[i] = src[i * stride], 8-byte words.
__global__ void copy_words_kernel(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, int count, long long stride) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[(long long)i * stride];
    __threadfence_system();
}
cudaError_t copy_words_launch(void* dst_mapped, const void* src, int count, int64_t stride, cudaStream_t st) {
    copy_words_kernel<<<1, 64, 0, st>>>(reinterpret_cast<unsigned long long*>(dst_mapped), reinterpret_cast<const unsigned long long*>(src), count,
                                         (long long)stride);
    return cudaGetLastError();
}

}  // namespace exb
