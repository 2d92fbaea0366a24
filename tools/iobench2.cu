// iobench2.cu -- page cache -> pinned ring -> GPU, pipelined at slice granularity (measurement tool, not product code).
//   iobench2 <file> [device]
// For slice sizes and thread counts: T workers memcpy slice k from a mapping of the file into a pinned ring slot, the
// main thread enqueues its H2D copy and the workers go on with slice k+1.  Small slices keep the freshly written lines in
// the CPU's last-level cache, so the DMA read may be served from there instead of DRAM.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

// persistent workers: run(fn, parts) executes fn(k) for k in [0, parts) on T threads (caller included)
struct Pool {
    int T;
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, done_cv;
    const std::function<void(int)>* fn = nullptr;
    int parts = 0, next = 0, left = 0;
    uint64_t gen = 0;
    bool stop = false;
    explicit Pool(int t) : T(t) {
        for (int i = 0; i < T - 1; i++) th.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
    void drain(std::unique_lock<std::mutex>& lk) {
        while (next < parts) {
            int k = next++;
            lk.unlock();
            (*fn)(k);
            lk.lock();
            if (--left == 0) done_cv.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            drain(lk);
        }
    }
    void run(const std::function<void(int)>& f, int n) {
        std::unique_lock<std::mutex> lk(mu);
        fn = &f;
        parts = n;
        next = 0;
        left = n;
        gen++;
        cv.notify_all();
        drain(lk);
        done_cv.wait(lk, [&] { return left == 0; });
    }
};

static void copy_temporal(uint8_t* d, const uint8_t* s, size_t n) {  // regular (cached) 64-byte stores
    size_t i = 0;
    for (; i + 64 <= n; i += 64) _mm512_storeu_si512((void*)(d + i), _mm512_loadu_si512((const void*)(s + i)));
    if (i < n) memcpy(d + i, s + i, n - i);
}
static void copy_nt(uint8_t* d, const uint8_t* s, size_t n) {  // streaming stores (d 64-byte aligned)
    size_t i = 0;
    for (; i + 64 <= n; i += 64) _mm512_stream_si512((__m512i*)(d + i), _mm512_loadu_si512((const void*)(s + i)));
    _mm_sfence();
    if (i < n) memcpy(d + i, s + i, n - i);
}

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    CK(cudaSetDevice(argc > 2 ? atoi(argv[2]) : 0));
    struct stat sb;
    if (stat(argv[1], &sb) != 0) return 1;
    const int64_t N = std::min<int64_t>(sb.st_size, 3ll << 30) & ~(int64_t)((64 << 20) - 1);
    int fd = open(argv[1], O_RDONLY);
    const uint8_t* m = (const uint8_t*)mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) return 1;
    const int64_t RING = 256ll << 20;
    uint8_t *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc(&h, (size_t)RING, cudaHostAllocPortable));
    CK(cudaMalloc(&d, (size_t)N));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    {  // touch the mapping once (page faults) so every configuration sees the same state
        volatile uint64_t acc = 0;
        for (int64_t i = 0; i < N; i += 4096) acc += m[i];
    }
    printf("input %.2f GB, ring %lld MiB, host threads %u\n", N / 1e9, (long long)(RING >> 20), std::thread::hardware_concurrency());
    for (int kind = 0; kind < 3; kind++) {
        const char* kname = kind == 0 ? "memcpy" : (kind == 1 ? "avx512 cached" : "avx512 streaming");
        for (int T : {8, 12, 16}) {
            if (getenv("IOB_QUICK") && T != 8) continue;
            Pool pool(T);
            for (int64_t slice : {2ll << 20, 4ll << 20, 8ll << 20, 16ll << 20, 64ll << 20}) {
                if (getenv("IOB_QUICK") && slice != (64ll << 20)) continue;
                const int slots = (int)(RING / slice);
                std::vector<cudaEvent_t> ev(slots);
                for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                double best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    const double t0 = now();
                    int64_t k = 0;
                    for (int64_t p = 0; p < N; p += slice, k++) {
                        const int slot = (int)(k % slots);
                        if (k >= slots) CK(cudaEventSynchronize(ev[slot]));  // the slot's previous DMA has drained
                        uint8_t* dst = h + (int64_t)slot * slice;
                        const uint8_t* src = m + p;
                        const int64_t len = std::min(slice, N - p);
                        const int parts = T;
                        std::function<void(int)> f = [&](int i) {
                            const int64_t lo = (len * i / parts) & ~(int64_t)63, hi = i + 1 == parts ? len : ((len * (i + 1) / parts) & ~(int64_t)63);
                            if (kind == 0) memcpy(dst + lo, src + lo, (size_t)(hi - lo));
                            else if (kind == 1) copy_temporal(dst + lo, src + lo, (size_t)(hi - lo));
                            else copy_nt(dst + lo, src + lo, (size_t)(hi - lo));
                        };
                        pool.run(f, parts);
                        CK(cudaMemcpyAsync(d + p, dst, (size_t)len, cudaMemcpyHostToDevice, s));
                        CK(cudaEventRecord(ev[slot], s));
                    }
                    CK(cudaStreamSynchronize(s));
                    best = std::min(best, now() - t0);
                }
                printf("%-17s %2d threads, slice %3lld MiB: %.1f GB/s\n", kname, T, (long long)(slice >> 20), N / 1e9 / best);
                fflush(stdout);
                for (auto& e : ev) cudaEventDestroy(e);
            }
        }
    }
    return 0;
}
