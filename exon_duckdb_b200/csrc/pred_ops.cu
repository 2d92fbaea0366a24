// pred_ops.cu -- row predicates pushed down through new_reader's `filters`
// text (the reference serialises DuckDB table filters with FilterToString,
// exon/src/exon/arrow_table_function/module.cpp:158-214, and DataFusion
// evaluates them; here they are evaluated on the device before any string is
// gathered, so filtered-out rows never cross PCIe).
#include "common.cuh"
#include "exon_b200_internal.h"
#include "x87div.h"

namespace exb {

// VARCHAR comparison = unsigned byte-wise lexicographic order (DuckDB and
// DataFusion agree for Utf8).  op: EXB_OP_*; 6 = IS NULL, 7 = IS NOT NULL.
__global__ void __launch_bounds__(256) str_pred_kernel(const uint8_t* __restrict__ buf, const int64_t* __restrict__ start,
                                                       const uint32_t* __restrict__ len, const uint8_t* __restrict__ valid, int64_t n_rows,
                                                       int op, const uint8_t* __restrict__ cst, int cst_len, uint8_t* __restrict__ pass) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        const bool is_valid = valid ? valid[i] != 0 : true;
        bool r;
        if (op == 6) r = !is_valid;
        else if (op == 7) r = is_valid;
        else if (!is_valid) r = false;  // NULL <op> const is NULL: the row is filtered out
        else {
            const uint8_t* p = buf + start[i];
            const int64_t l = len[i];
            const int64_t m = l < cst_len ? l : cst_len;
            int c = 0;
            for (int64_t k = 0; k < m; k++) {
                int a = p[k], b = cst[k];
                if (a != b) {
                    c = a < b ? -1 : 1;
                    break;
                }
            }
            if (c == 0) c = l < cst_len ? -1 : (l > cst_len ? 1 : 0);
            switch (op) {
            case EXB_OP_GT: r = c > 0; break;
            case EXB_OP_GE: r = c >= 0; break;
            case EXB_OP_LT: r = c < 0; break;
            case EXB_OP_LE: r = c <= 0; break;
            case EXB_OP_EQ: r = c == 0; break;
            default: r = c != 0; break;
            }
        }
        pass[i] = r ? 1 : 0;
    }
}

// a = a AND b (mode 0) / a OR b (mode 1)
__global__ void pass_combine_kernel(uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int64_t n, int mode) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        a[i] = mode == 0 ? (a[i] & b[i]) : (a[i] | b[i]);
}

// FASTA numeric predicates from the scan's prefix arrays: field EXB_P_GC_CONTENT or EXB_P_SEQ_LEN
__global__ void fasta_num_pred_kernel(const int64_t* __restrict__ seq_off, const int64_t* __restrict__ gc_prefix, int64_t n_rows, int field,
                                      int op, double value, uint8_t* __restrict__ pass) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const long long len = seq_off[r + 1] - seq_off[r];
        double v;
        if (field == EXB_P_GC_CONTENT) {
            const long long gc = gc_prefix[r + 1] - gc_prefix[r];
            v = (double)(len == 0 ? 0.0f : __fdiv_rn(__ll2float_rn(gc), __ll2float_rn(len)));
        } else {
            v = (double)len;
        }
        pass[r] = exb_cmp(v, op, value) ? 1 : 0;
    }
}

// lens / starts of the sequence column of FASTA rows (optionally a selection) from the prefix array
__global__ void fasta_seq_ranges_kernel(const int64_t* __restrict__ seq_off, const int64_t* __restrict__ sel, int64_t n_rows,
                                        int64_t* __restrict__ start, uint32_t* __restrict__ len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = sel ? sel[i] : i;
        start[i] = seq_off[r];
        len[i] = (uint32_t)(seq_off[r + 1] - seq_off[r]);
    }
}

// out[i] = in[sel[i]] for the small per-row arrays (uint32 lens, int64 starts, uint8 validity)
template <typename T>
__global__ void take_kernel(const T* __restrict__ in, const int64_t* __restrict__ sel, int64_t n_rows, T* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[sel[i]];
}

static int blocks_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return b < 1 ? 1 : (int)b;
}

cudaError_t str_pred_launch(const uint8_t* buf, const int64_t* start, const uint32_t* len, const uint8_t* valid, int64_t n_rows, int op,
                            const uint8_t* cst, int cst_len, uint8_t* pass, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    str_pred_kernel<<<blocks_for(n_rows), 256, 0, st>>>(buf, start, len, valid, n_rows, op, cst, cst_len, pass);
    return cudaGetLastError();
}
cudaError_t pass_combine_launch(uint8_t* a, const uint8_t* b, int64_t n, int mode, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    pass_combine_kernel<<<blocks_for(n), 256, 0, st>>>(a, b, n, mode);
    return cudaGetLastError();
}
cudaError_t fasta_num_pred_launch(const int64_t* seq_off, const int64_t* gc_prefix, int64_t n_rows, int field, int op, double value,
                                  uint8_t* pass, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    fasta_num_pred_kernel<<<blocks_for(n_rows), 256, 0, st>>>(seq_off, gc_prefix, n_rows, field, op, value, pass);
    return cudaGetLastError();
}
cudaError_t fasta_seq_ranges_launch(const int64_t* seq_off, const int64_t* sel, int64_t n_rows, int64_t* start, uint32_t* len,
                                    cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    fasta_seq_ranges_kernel<<<blocks_for(n_rows), 256, 0, st>>>(seq_off, sel, n_rows, start, len);
    return cudaGetLastError();
}
cudaError_t take_u32_launch(const uint32_t* in, const int64_t* sel, int64_t n, uint32_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    take_kernel<uint32_t><<<blocks_for(n), 256, 0, st>>>(in, sel, n, out);
    return cudaGetLastError();
}
cudaError_t take_i64_launch(const int64_t* in, const int64_t* sel, int64_t n, int64_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    take_kernel<int64_t><<<blocks_for(n), 256, 0, st>>>(in, sel, n, out);
    return cudaGetLastError();
}
cudaError_t take_u8_launch(const uint8_t* in, const int64_t* sel, int64_t n, uint8_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    take_kernel<uint8_t><<<blocks_for(n), 256, 0, st>>>(in, sel, n, out);
    return cudaGetLastError();
}

}  // namespace exb
