// synth.cu -- libexb_synth.so: the deterministic FASTA / FASTQ generators of SURVEY 8(d) (C1-C5), on the host and on
// the device.  Test and bench tooling only; the product library (exon_duckdb_b200/libexon_b200.so) does not contain it.
#include <cuda_runtime.h>
#include <stdio.h>

#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include "exb_synth.h"
#include "gen_common.h"

namespace {
__global__ void gen_sizes_kernel(exb_gen_params p, long long* sizes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_records; i += (int64_t)gridDim.x * blockDim.x)
        sizes[i] = (long long)exb_gen_record(&p, (uint64_t)(p.first_record + i), nullptr);
}
__global__ void gen_fill_kernel(exb_gen_params p, const long long* off, uint8_t* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_records; i += (int64_t)gridDim.x * blockDim.x)
        exb_gen_record(&p, (uint64_t)(p.first_record + i), out + off[i]);
}
int fail(cudaError_t e, const char* what) {
    fprintf(stderr, "exb_synth: %s: %s\n", what, cudaGetErrorString(e));
    return -1;
}
}  // namespace

extern "C" {

int64_t exb_gen_size(const exb_gen_params* p) {
    int64_t tot = 0;
    for (int64_t i = 0; i < p->n_records; i++) tot += exb_gen_record(p, (uint64_t)(p->first_record + i), nullptr);
    return tot;
}

int exb_gen_host(const exb_gen_params* p, void* out, int64_t cap) {
    uint8_t* o = reinterpret_cast<uint8_t*>(out);
    int64_t at = 0;
    for (int64_t i = 0; i < p->n_records; i++) {
        int64_t sz = exb_gen_record(p, (uint64_t)(p->first_record + i), nullptr);
        if (at + sz > cap) return -4;
        exb_gen_record(p, (uint64_t)(p->first_record + i), o + at);
        at += sz;
    }
    return 0;
}

// Size of the text computed on the DEVICE (the host loop of exb_gen_size takes ~30 s for the 287 M records of C5).
int exb_gen_size_device(const exb_gen_params* p, int64_t* size_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = p->n_records;
    *size_out = 0;
    if (n == 0) return 0;
    long long *sizes = nullptr, *total = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaMalloc(&sizes, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&total, 8);
    if (e == cudaSuccess) e = cub::DeviceReduce::Sum(nullptr, tmp_bytes, sizes, total, (int)n, st);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes + 16);
    int rc = 0;
    if (e != cudaSuccess) rc = fail(e, "scratch");
    if (!rc) {
        int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
        gen_sizes_kernel<<<blocks, 256, 0, st>>>(*p, sizes);
        e = cub::DeviceReduce::Sum(tmp, tmp_bytes, sizes, total, (int)n, st);
        long long t = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&t, total, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(e, "sizes");
        *size_out = t;
    }
    cudaFree(sizes);
    cudaFree(total);
    cudaFree(tmp);
    return rc;
}

int exb_gen_device(const exb_gen_params* p, void* d_out, int64_t cap, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = p->n_records;
    if (n == 0) return 0;
    long long *sizes = nullptr, *off = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaMalloc(&sizes, (size_t)(n + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&off, (size_t)(n + 1) * 8);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sizes, off, (int)(n + 1), st);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes + 16);
    int rc = 0;
    if (e != cudaSuccess) rc = fail(e, "scratch");
    if (!rc) {
        int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
        cudaMemsetAsync(sizes + n, 0, 8, st);
        gen_sizes_kernel<<<blocks, 256, 0, st>>>(*p, sizes);
        e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, sizes, off, (int)(n + 1), st);
        long long total = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&total, off + n, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(e, "sizes");
        else if (total > cap) rc = -4;
        else {
            // few long records (genomes): one thread each is still the simplest correct mapping
            int fblocks = (int)((n + 63) / 64 < 148 * 64 ? (n + 63) / 64 : 148 * 64);
            gen_fill_kernel<<<fblocks, 64, 0, st>>>(*p, off, reinterpret_cast<uint8_t*>(d_out));
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = fail(e, "fill");
        }
    }
    cudaFree(sizes);
    cudaFree(off);
    cudaFree(tmp);
    return rc;
}

}  // extern "C"
