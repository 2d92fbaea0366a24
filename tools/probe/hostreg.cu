// Can the page cache be registered for DMA?  Tries cudaHostRegister on a tmpfs file mapping in several flavours and
// times registration, H2D from the registered mapping and unregistration.  usage: hostreg <file>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdint.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "/dev/shm/hostreg.bin";
    int fd = open(path, O_RDWR);
    if (fd < 0) { perror("open"); return 1; }
    struct stat sb; fstat(fd, &sb);
    int64_t N = sb.st_size & ~((int64_t)(2 << 20) - 1);
    if (N > (4ll << 30)) N = 4ll << 30;
    int ro = 0; cudaDeviceGetAttribute(&ro, cudaDevAttrHostRegisterReadOnlySupported, 0);
    int hr = 0; cudaDeviceGetAttribute(&hr, cudaDevAttrHostRegisterSupported, 0);
    printf("file %.2f GB; HostRegisterSupported %d, ReadOnlySupported %d\n", N / 1e9, hr, ro);
    uint8_t* d; cudaMalloc(&d, (size_t)N);
    cudaStream_t s; cudaStreamCreate(&s);
    struct V { const char* name; int prot; int flags; unsigned reg; };
    V vs[] = {
        {"PROT_READ  MAP_SHARED            ReadOnly", PROT_READ, MAP_SHARED, cudaHostRegisterReadOnly},
        {"PROT_RW    MAP_SHARED            Default ", PROT_READ | PROT_WRITE, MAP_SHARED, cudaHostRegisterDefault},
        {"PROT_RW    MAP_SHARED|POPULATE   Default ", PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, cudaHostRegisterDefault},
        {"PROT_RW    MAP_SHARED|POPULATE   ReadOnly", PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, cudaHostRegisterReadOnly},
        {"PROT_RW    MAP_PRIVATE|POPULATE  Default ", PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_POPULATE, cudaHostRegisterDefault},
        {"PROT_READ  MAP_PRIVATE|POPULATE  ReadOnly", PROT_READ, MAP_PRIVATE | MAP_POPULATE, cudaHostRegisterReadOnly},
    };
    for (auto& v : vs) {
        void* m = mmap(nullptr, (size_t)N, v.prot, v.flags, fd, 0);
        if (m == MAP_FAILED) { printf("%s: mmap failed\n", v.name); continue; }
        double t0 = now();
        cudaError_t e = cudaHostRegister(m, (size_t)N, v.reg);
        double t1 = now();
        if (e != cudaSuccess) {
            printf("%s: cudaHostRegister -> %s\n", v.name, cudaGetErrorString(e));
            cudaGetLastError();
            munmap(m, (size_t)N);
            continue;
        }
        cudaMemcpyAsync(d, m, (size_t)N, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);
        double t2 = now();
        cudaMemcpyAsync(d, m, (size_t)N, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s);
        double t3 = now();
        cudaHostUnregister(m);
        double t4 = now();
        printf("%s: register %.1f GB/s (%.0f ms), H2D %.1f / %.1f GB/s, unregister %.0f ms\n", v.name, N / 1e9 / (t1 - t0), (t1 - t0) * 1e3,
               N / 1e9 / (t2 - t1), N / 1e9 / (t3 - t2), (t4 - t3) * 1e3);
        // registration in 64 MiB pieces by 8 threads
        {
            const int64_t B = 64ll << 20; const int T = 8;
            std::vector<std::thread> th; double a0 = now();
            for (int k = 0; k < T; k++) th.emplace_back([=] { cudaSetDevice(0); for (int64_t p = k * B; p < N; p += T * B) cudaHostRegister((uint8_t*)m + p, (size_t)std::min(B, N - p), v.reg); });
            for (auto& t : th) t.join();
            double a1 = now();
            printf("      registered again in 64 MiB pieces by %d threads: %.1f GB/s\n", T, N / 1e9 / (a1 - a0));
            for (int64_t p = 0; p < N; p += B) cudaHostUnregister((uint8_t*)m + p);
        }
        munmap(m, (size_t)N);
    }
    return 0;
}
