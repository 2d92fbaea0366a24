// engine.cu -- host-buffer (end-to-end) FASTQ aggregate path.
//
// An exb_engine streams a FASTQ image that lives in HOST memory through the
// GPU: chunked cudaMemcpyAsync on a copy stream into ONE device buffer, each
// chunk scanned on a compute stream as soon as its copy has landed
// (exb_fastq_scan chained through d_prev_workspace, so records may straddle
// chunk edges freely), then one filter/aggregate pass over the per-record
// arrays.  The copy of chunk k+1 overlaps the scan of chunk k; the kernels run
// two orders of magnitude faster than PCIe, so the wall time is the H2D time.
// Device buffers, streams and events are owned by the engine and reused
// across calls (no allocation in the steady state).
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

#include "exon_b200_internal.h"

namespace exb {
int set_err(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
}  // namespace exb
using namespace exb;

struct exb_engine {
    int device = 0;
    int64_t chunk_bytes = 0;
    cudaStream_t sc = nullptr, sk = nullptr;
    void* d_buf = nullptr;
    int64_t buf_cap = 0;
    void* d_ws[2] = {nullptr, nullptr};
    int64_t ws_bytes = 0;
    void* d_arr[4] = {nullptr, nullptr, nullptr, nullptr};  // seq_len, gc, qual_len, qsum
    int64_t rec_cap = 0;
    void* d_agg = nullptr;
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_free[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> events;
};

static int ensure_dev(void** p, int64_t* cap, int64_t need, const char* what) {
    if (*p && *cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    cudaError_t e = cudaMalloc(p, (size_t)(need > 0 ? need : 16));
    if (e != cudaSuccess) return cuda_fail(e, what);
    *cap = need;
    return 0;
}

extern "C" {

int exb_engine_create(int device, int64_t chunk_bytes, exb_engine** out) {
    if (!out) return set_err(EXB_ERR_ARG, "exb_engine_create: null out");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    exb_engine* g = new exb_engine();
    g->device = device;
    if (chunk_bytes <= 0) chunk_bytes = 64ll << 20;
    g->chunk_bytes = (chunk_bytes + 16383) & ~16383ll;  // whole tiles so chained ranges stay aligned
    if ((e = cudaStreamCreateWithFlags(&g->sc, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&g->sk, cudaStreamNonBlocking)) != cudaSuccess) {
        delete g;
        return cuda_fail(e, "cudaStreamCreate");
    }
    g->ws_bytes = exb_scan_workspace_bytes(g->chunk_bytes);
    for (int i = 0; i < 2; i++)
        if ((e = cudaMalloc(&g->d_ws[i], (size_t)g->ws_bytes)) != cudaSuccess) {
            delete g;
            return cuda_fail(e, "cudaMalloc(workspace)");
        }
    if ((e = cudaMalloc(&g->d_agg, 64)) != cudaSuccess) {
        delete g;
        return cuda_fail(e, "cudaMalloc(agg)");
    }
    *out = g;
    return 0;
}

void exb_engine_destroy(exb_engine* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->sc) cudaStreamSynchronize(g->sc);
    if (g->sk) cudaStreamSynchronize(g->sk);
    for (cudaEvent_t ev : g->events) cudaEventDestroy(ev);
    for (int i = 0; i < 2; i++) {
        if (g->stage_free[i]) cudaEventDestroy(g->stage_free[i]);
        if (g->stage[i]) cudaFreeHost(g->stage[i]);
        if (g->d_ws[i]) cudaFree(g->d_ws[i]);
    }
    for (int i = 0; i < 4; i++)
        if (g->d_arr[i]) cudaFree(g->d_arr[i]);
    if (g->d_agg) cudaFree(g->d_agg);
    if (g->d_buf) cudaFree(g->d_buf);
    if (g->sc) cudaStreamDestroy(g->sc);
    if (g->sk) cudaStreamDestroy(g->sk);
    delete g;
}

int exb_engine_fastq_count(exb_engine* g, const void* host_buf, int64_t n, const exb_predicate* preds, int n_preds, int64_t* agg_out,
                           exb_scan_result* res_out) {
    if (!g || (!host_buf && n > 0) || n < 0 || !agg_out) return set_err(EXB_ERR_ARG, "exb_engine_fastq_count: bad arguments");
    cudaError_t e = cudaSetDevice(g->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const int flags = EXB_F_SEQ | EXB_F_QUAL;  // the aggregates report sums of both families

    cudaPointerAttributes attr;
    bool pinned = false;
    if (n > 0 && cudaPointerGetAttributes(&attr, host_buf) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    cudaGetLastError();

    int rc = ensure_dev(&g->d_buf, &g->buf_cap, n + 64, "cudaMalloc(input)");
    if (rc) return rc;
    if (!pinned && n > 0)
        for (int i = 0; i < 2; i++)
            if (!g->stage[i]) {
                if ((e = cudaHostAlloc(&g->stage[i], (size_t)g->chunk_bytes, cudaHostAllocDefault)) != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
                cudaEventCreateWithFlags(&g->stage_free[i], cudaEventDisableTiming);
            }
    const int64_t chunk = g->chunk_bytes;
    const int64_t n_chunks = n == 0 ? 1 : (n + chunk - 1) / chunk;
    while ((int64_t)g->events.size() < n_chunks) {
        cudaEvent_t ev;
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        g->events.push_back(ev);
    }
    // the estimate follows the CURRENT input (a capacity kept from an earlier, smaller file would overflow first and redo the pass)
    int64_t want_cap = g->rec_cap > n / 32 + 4096 ? g->rec_cap : n / 32 + 4096;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (g->rec_cap < want_cap) {
            for (int i = 0; i < 4; i++) {
                if (g->d_arr[i]) cudaFree(g->d_arr[i]);
                g->d_arr[i] = nullptr;
                if ((e = cudaMalloc(&g->d_arr[i], (size_t)want_cap * 4)) != cudaSuccess) {
                    g->rec_cap = 0;
                    return cuda_fail(e, "cudaMalloc(per-record arrays)");
                }
            }
            g->rec_cap = want_cap;
        }
        const void* prev_ws = nullptr;
        for (int64_t c = 0; c < n_chunks; c++) {
            const int64_t off = c * chunk;
            const int64_t len = (n - off < chunk) ? (n - off) : chunk;
            if (len > 0) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(host_buf) + off;
                if (!pinned) {  // pageable source: bounce through the engine's pinned staging pair
                    const int s = (int)(c & 1);
                    if (c >= 2) cudaEventSynchronize(g->stage_free[s]);
                    memcpy(g->stage[s], src, (size_t)len);
                    e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(g->d_buf) + off, g->stage[s], (size_t)len, cudaMemcpyHostToDevice, g->sc);
                    cudaEventRecord(g->stage_free[s], g->sc);
                } else {
                    e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(g->d_buf) + off, src, (size_t)len, cudaMemcpyHostToDevice, g->sc);
                }
                if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
            }
            cudaEventRecord(g->events[c], g->sc);
            cudaStreamWaitEvent(g->sk, g->events[c], 0);
            void* ws = g->d_ws[c & 1];
            rc = exb_fastq_scan(g->d_buf, off, off + len, c == n_chunks - 1, prev_ws, ~0ull, flags, nullptr, 0, 0, (uint32_t*)g->d_arr[0],
                                (uint32_t*)g->d_arr[1], (uint32_t*)g->d_arr[2], (int32_t*)g->d_arr[3], g->rec_cap, ws, g->ws_bytes, g->sk);
            if (rc) return rc;
            prev_ws = ws;
        }
        exb_scan_result res;
        rc = exb_scan_result_fetch(prev_ws, &res, g->sk);
        if (rc) return rc;
        if (res.overflow && attempt == 0) {  // denser records than estimated: redo with the hard upper bounds
            want_cap = n / 4 + 16;
            cudaStreamSynchronize(g->sc);
            const int64_t ws_need = exb_fastq_workspace_bytes(g->chunk_bytes, g->chunk_bytes + 1);
            for (int i = 0; i < 2; i++) {
                cudaFree(g->d_ws[i]);
                g->d_ws[i] = nullptr;
                if ((e = cudaMalloc(&g->d_ws[i], (size_t)ws_need)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)");
            }
            g->ws_bytes = ws_need;
            continue;
        }
        if (res_out) *res_out = res;
        if (res.err_pos != ~0ull) return set_err(EXB_ERR_FORMAT, "malformed FASTQ record at byte %llu", (unsigned long long)res.err_pos);
        if (res.total_lines % 4 != 0) return set_err(EXB_ERR_FORMAT, "truncated FASTQ record: %llu lines", (unsigned long long)res.total_lines);
        if (res.overflow) return set_err(EXB_ERR_CAPACITY, "per-record capacity exceeded");
        const int64_t n_rec = (int64_t)(res.total_lines / 4);
        rc = exb_fastq_filter((uint32_t*)g->d_arr[0], (uint32_t*)g->d_arr[1], (uint32_t*)g->d_arr[2], (int32_t*)g->d_arr[3], n_rec, preds,
                              n_preds, nullptr, (int64_t*)g->d_agg, nullptr, g->sk);
        if (rc) return rc;
        int64_t agg[8];
        e = cudaMemcpyAsync(agg, g->d_agg, sizeof(agg), cudaMemcpyDeviceToHost, g->sk);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g->sk);
        if (e != cudaSuccess) return cuda_fail(e, "fetch aggregates");
        for (int i = 0; i < 5; i++) agg_out[i] = agg[i];
        agg_out[5] = n_rec;
        return 0;
    }
    return set_err(EXB_ERR_CAPACITY, "per-record capacity exceeded");
}

int exb_fastq_count_host(const void* host_buf, int64_t n, const exb_predicate* preds, int n_preds, int64_t chunk_bytes, int device,
                         int64_t* agg_out, exb_scan_result* res_out) {
    exb_engine* g = nullptr;
    int rc = exb_engine_create(device, chunk_bytes, &g);
    if (rc) return rc;
    rc = exb_engine_fastq_count(g, host_buf, n, preds, n_preds, agg_out, res_out);
    exb_engine_destroy(g);
    return rc;
}

void* exb_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)(bytes > 0 ? bytes : 16), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void exb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"

// ------------------------------------------------------------------ scalar functions over host columns
// One context per calling thread (DuckDB calls scalar functions concurrently from its worker threads):
// a stream, growable device buffers and pinned staging, kept for the life of the thread.
namespace {
struct HostCtx {
    cudaStream_t st = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_off = nullptr;
    void* d_flag = nullptr;
    void* h_stage = nullptr;
    int64_t in_cap = 0, out_cap = 0, off_cap = 0, stage_cap = 0;
    bool ok = false;
    ~HostCtx() {
        if (d_in) cudaFree(d_in);
        if (d_out) cudaFree(d_out);
        if (d_off) cudaFree(d_off);
        if (d_flag) cudaFree(d_flag);
        if (h_stage) cudaFreeHost(h_stage);
        if (st) cudaStreamDestroy(st);
    }
    int init() {
        if (ok) return 0;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            cudaGetLastError();
            return set_err(EXB_ERR_CUDA, "no CUDA device: the exon_b200 scalar functions have no CPU fallback");
        }
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
        if ((e = cudaMalloc(&d_flag, 16)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        ok = true;
        return 0;
    }
    int grow(void** p, int64_t* cap, int64_t need, const char* what) {
        if (*p && *cap >= need) return 0;
        if (*p) cudaFree(*p);
        *p = nullptr;
        const int64_t want = need + need / 2 + 4096;
        cudaError_t e = cudaMalloc(p, (size_t)want);
        if (e != cudaSuccess) {
            *cap = 0;
            return cuda_fail(e, what);
        }
        *cap = want;
        return 0;
    }
    int grow_stage(int64_t need) {
        if (h_stage && stage_cap >= need) return 0;
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        const int64_t want = need + need / 2 + 4096;
        cudaError_t e = cudaHostAlloc(&h_stage, (size_t)want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            stage_cap = 0;
            return cuda_fail(e, "cudaHostAlloc(staging)");
        }
        stage_cap = want;
        return 0;
    }
};
thread_local HostCtx g_host;
}  // namespace

extern "C" {

int exb_gc_content_host(const int64_t* offsets, const uint8_t* data, int64_t n_rows, float* out) {
    if (n_rows < 0 || (n_rows > 0 && (!offsets || !out))) return set_err(EXB_ERR_ARG, "exb_gc_content_host: bad arguments");
    if (n_rows == 0) return 0;
    HostCtx& c = g_host;
    int rc = c.init();
    if (rc) return rc;
    const int64_t base = offsets[0], nb = offsets[n_rows] - base;
    if ((rc = c.grow(&c.d_in, &c.in_cap, nb + 64, "cudaMalloc(in)")) || (rc = c.grow(&c.d_off, &c.off_cap, (n_rows + 1) * 8, "cudaMalloc(off)")) ||
        (rc = c.grow(&c.d_out, &c.out_cap, n_rows * 4, "cudaMalloc(out)")))
        return rc;
    cudaError_t e = cudaMemcpyAsync(c.d_off, offsets, (size_t)(n_rows + 1) * 8, cudaMemcpyHostToDevice, c.st);
    if (e == cudaSuccess && nb > 0) e = cudaMemcpyAsync(c.d_in, data + base, (size_t)nb, cudaMemcpyHostToDevice, c.st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    // the offsets keep their base: the kernel addresses d_data[off], so hand it a pointer shifted by -base
    rc = exb_gc_content((const int64_t*)c.d_off, (const uint8_t*)c.d_in - base, n_rows, (float*)c.d_out, c.st);
    if (rc) return rc;
    e = cudaMemcpyAsync(out, c.d_out, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, c.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
    if (e != cudaSuccess) return cuda_fail(e, "gc_content D2H");
    return 0;
}

int exb_seq_map_host(const uint8_t* data, int64_t n_bytes, int mode, uint8_t* out, int64_t* bad_pos) {
    if (n_bytes < 0 || (n_bytes > 0 && (!data || !out)) || !bad_pos) return set_err(EXB_ERR_ARG, "exb_seq_map_host: bad arguments");
    *bad_pos = -1;
    if (n_bytes == 0) return 0;
    HostCtx& c = g_host;
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.grow(&c.d_in, &c.in_cap, n_bytes + 64, "cudaMalloc(in)")) || (rc = c.grow(&c.d_out, &c.out_cap, n_bytes + 64, "cudaMalloc(out)")))
        return rc;
    cudaError_t e = cudaMemcpyAsync(c.d_in, data, (size_t)n_bytes, cudaMemcpyHostToDevice, c.st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    rc = exb_seq_map((const uint8_t*)c.d_in, n_bytes, mode, (uint8_t*)c.d_out, (uint64_t*)c.d_flag, c.st);
    if (rc) return rc;
    uint64_t bad = 0;
    e = cudaMemcpyAsync(out, c.d_out, (size_t)n_bytes, cudaMemcpyDeviceToHost, c.st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, c.d_flag, 8, cudaMemcpyDeviceToHost, c.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
    if (e != cudaSuccess) return cuda_fail(e, "seq_map D2H");
    *bad_pos = bad == ~0ull ? -1 : (int64_t)bad;
    return 0;
}

int exb_translate_host(const int64_t* offsets, const uint8_t* data, int64_t n_rows, uint8_t* out, int64_t* status) {
    if (n_rows < 0 || !status || (n_rows > 0 && (!offsets || !out))) return set_err(EXB_ERR_ARG, "exb_translate_host: bad arguments");
    status[0] = status[1] = -1;
    if (n_rows == 0) return 0;
    const int64_t nb = offsets[n_rows] - offsets[0];
    HostCtx& c = g_host;
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.grow(&c.d_in, &c.in_cap, nb + 64, "cudaMalloc(in)")) || (rc = c.grow(&c.d_off, &c.off_cap, (n_rows + 1) * 8 + 16, "cudaMalloc(off)")) ||
        (rc = c.grow(&c.d_out, &c.out_cap, nb / 3 + 64, "cudaMalloc(out)")))
        return rc;
    cudaError_t e = cudaMemcpyAsync(c.d_off, offsets, (size_t)(n_rows + 1) * 8, cudaMemcpyHostToDevice, c.st);
    if (e == cudaSuccess && nb > 0) e = cudaMemcpyAsync(c.d_in, data + offsets[0], (size_t)nb, cudaMemcpyHostToDevice, c.st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    // the offsets keep their base: hand the kernel a data pointer shifted by -offsets[0]
    rc = exb_translate((const int64_t*)c.d_off, (const uint8_t*)c.d_in - offsets[0], n_rows, (uint8_t*)c.d_out, (int64_t*)c.d_flag, c.st);
    if (rc) return rc;
    e = cudaMemcpyAsync(status, c.d_flag, 16, cudaMemcpyDeviceToHost, c.st);
    if (e == cudaSuccess && nb >= 3) e = cudaMemcpyAsync(out, c.d_out, (size_t)(nb / 3), cudaMemcpyDeviceToHost, c.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
    if (e != cudaSuccess) return cuda_fail(e, "translate D2H");
    return 0;
}

int exb_quality_decode_host(const uint8_t* data, int64_t n_bytes, int32_t* out) {
    if (n_bytes < 0 || (n_bytes > 0 && (!data || !out))) return set_err(EXB_ERR_ARG, "exb_quality_decode_host: bad arguments");
    if (n_bytes == 0) return 0;
    HostCtx& c = g_host;
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.grow(&c.d_in, &c.in_cap, n_bytes + 64, "cudaMalloc(in)")) || (rc = c.grow(&c.d_out, &c.out_cap, n_bytes * 4 + 64, "cudaMalloc(out)")))
        return rc;
    cudaError_t e = cudaMemcpyAsync(c.d_in, data, (size_t)n_bytes, cudaMemcpyHostToDevice, c.st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
    rc = exb_quality_decode((const uint8_t*)c.d_in, n_bytes, (int32_t*)c.d_out, c.st);
    if (rc) return rc;
    e = cudaMemcpyAsync(out, c.d_out, (size_t)n_bytes * 4, cudaMemcpyDeviceToHost, c.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
    if (e != cudaSuccess) return cuda_fail(e, "quality_decode D2H");
    return 0;
}

}  // extern "C"
