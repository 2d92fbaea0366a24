// iobench.cu -- how fast can file bytes reach the GPU on this box?  (measurement tool, not product code)
//   iobench <file> [device]
// Reports: pinned H2D peak; pread into pinned memory with T threads (page cache -> pinned = a kernel memcpy);
// mmap + memcpy; mmap + cudaHostRegister (DMA straight from the page cache); plus the host topology the
// reader's NUMA binding depends on.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));  \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

static void pread_par(int fd, uint8_t* dst, int64_t pos, int64_t n, int T) {
    std::vector<std::thread> th;
    for (int k = 0; k < T; k++)
        th.emplace_back([=] {
            int64_t lo = n * k / T, hi = n * (k + 1) / T;
            while (lo < hi) {
                ssize_t r = pread(fd, dst + lo, (size_t)(hi - lo), (off_t)(pos + lo));
                if (r <= 0) break;
                lo += r;
            }
        });
    for (auto& t : th) t.join();
}

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    const char* path = argv[1];
    int dev = argc > 2 ? atoi(argv[2]) : 0;
    CK(cudaSetDevice(dev));
    struct stat sb;
    if (stat(path, &sb) != 0) { perror("stat"); return 1; }
    const int64_t fsz = sb.st_size;
    const int64_t N = std::min<int64_t>(fsz, 2ll << 30);
    printf("file %s: %.2f GB, using %.2f GB; host threads %u\n", path, fsz / 1e9, N / 1e9, std::thread::hardware_concurrency());
    uint8_t *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc(&h, (size_t)N, cudaHostAllocDefault));
    CK(cudaMalloc(&d, (size_t)N));
    int fd = open(path, O_RDONLY);
    pread_par(fd, h, 0, N, 8);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    // 1. pinned H2D peak
    for (int rep = 0; rep < 3; rep++) {
        double t0 = now();
        CK(cudaMemcpyAsync(d, h, (size_t)N, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        double t = now() - t0;
        printf("H2D pinned %.2f GB in %.1f ms = %.1f GB/s\n", N / 1e9, t * 1e3, N / 1e9 / t);
    }
    {
        double t0 = now();
        CK(cudaMemcpyAsync(h, d, (size_t)N, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        double t = now() - t0;
        printf("D2H pinned = %.1f GB/s\n", N / 1e9 / t);
    }
    // 2. pread into pinned, T threads, 64 MiB blocks
    for (int T : {1, 2, 4, 8, 12, 16, 24, 32, 48}) {
        const int64_t B = 64ll << 20;
        double t0 = now();
        for (int64_t p = 0; p < N; p += B) pread_par(fd, h + p, p, std::min(B, N - p), T);
        double t = now() - t0;
        printf("pread -> pinned, %2d threads per 64 MiB block: %.1f GB/s\n", T, N / 1e9 / t);
    }
    // 2b. whole-range pread (no per-block join)
    for (int T : {8, 16, 32}) {
        double t0 = now();
        pread_par(fd, h, 0, N, T);
        double t = now() - t0;
        printf("pread -> pinned, %2d threads over the whole range: %.1f GB/s\n", T, N / 1e9 / t);
    }
    // 2c. pread overlapped with H2D of the previous block (what the reader does), 16 threads
    for (int T : {8, 16, 32}) {
        const int64_t B = 64ll << 20;
        double t0 = now();
        for (int64_t p = 0; p < N; p += B) {
            int64_t len = std::min(B, N - p);
            pread_par(fd, h + p, p, len, T);
            CK(cudaMemcpyAsync(d + p, h + p, (size_t)len, cudaMemcpyHostToDevice, s));
        }
        CK(cudaStreamSynchronize(s));
        double t = now() - t0;
        printf("pread(%2d thr) + async H2D pipelined: %.1f GB/s\n", T, N / 1e9 / t);
    }
    // 3. mmap + memcpy to pinned
    void* m = mmap(nullptr, (size_t)N, PROT_READ, MAP_SHARED, fd, 0);
    if (m != MAP_FAILED) {
        for (int T : {8, 16, 32}) {
            std::vector<std::thread> th;
            double t0 = now();
            for (int k = 0; k < T; k++)
                th.emplace_back([=] {
                    int64_t lo = N * k / T, hi = N * (k + 1) / T;
                    memcpy(h + lo, (uint8_t*)m + lo, (size_t)(hi - lo));
                });
            for (auto& t : th) t.join();
            double t = now() - t0;
            printf("mmap memcpy -> pinned, %2d threads: %.1f GB/s\n", T, N / 1e9 / t);
        }
        // 4. cudaHostRegister on the mapping, block by block, then H2D from it
        const int64_t B = 64ll << 20;
        double t_reg = 0, t_cp = 0, t_unreg = 0;
        int64_t done = 0;
        bool ok = true;
        for (int64_t p = 0; p < N && ok; p += B) {
            int64_t len = std::min(B, N - p);
            double t0 = now();
            cudaError_t e = cudaHostRegister((uint8_t*)m + p, (size_t)len, cudaHostRegisterReadOnly);
            if (e != cudaSuccess) {
                cudaGetLastError();
                e = cudaHostRegister((uint8_t*)m + p, (size_t)len, cudaHostRegisterDefault);
            }
            if (e != cudaSuccess) {
                printf("cudaHostRegister on a file mapping failed: %s\n", cudaGetErrorString(e));
                cudaGetLastError();
                ok = false;
                break;
            }
            double t1 = now();
            CK(cudaMemcpyAsync(d + p, (uint8_t*)m + p, (size_t)len, cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
            double t2 = now();
            CK(cudaHostUnregister((uint8_t*)m + p));
            double t3 = now();
            t_reg += t1 - t0;
            t_cp += t2 - t1;
            t_unreg += t3 - t2;
            done += len;
        }
        if (ok)
            printf("mmap + cudaHostRegister per 64 MiB: register %.1f GB/s, H2D %.1f GB/s, unregister %.1f GB/s, serial total %.1f GB/s\n",
                   done / 1e9 / t_reg, done / 1e9 / t_cp, done / 1e9 / t_unreg, done / 1e9 / (t_reg + t_cp + t_unreg));
        // 5. plain cudaMemcpy from the pageable mapping
        double t0 = now();
        CK(cudaMemcpy(d, m, (size_t)N, cudaMemcpyHostToDevice));
        double t = now() - t0;
        printf("cudaMemcpy from the pageable mapping: %.1f GB/s\n", N / 1e9 / t);
        munmap(m, (size_t)N);
    }
    close(fd);
    return 0;
}
