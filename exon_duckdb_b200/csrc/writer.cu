// writer.cu -- C ABI of the FASTQ / FASTA writers (the kernels are in writer_ops.cu).
//
//   exb_fastq_format / exb_fasta_format / exb_format_finish   device layer: columns in HBM -> file image in HBM
//   exb_writer_open / _append / _close                        host layer behind `COPY ... TO (FORMAT 'fastq' | 'fasta')`:
//        host columns are staged in pinned memory, shipped once per ~64 MiB, formatted on the device, read back and
//        written (plain, gzip through zlib, zstd through the system's libzstd bound at run time -- the codecs the
//        reference's copy tests name: test_fastq_copy.test, test_fasta_copy.test).
// Reference interface replaced: FastqFunctions::GetFastqCopyFunction (exon/include/exon/fastq_functions/module.hpp:30)
// and its FASTA twin, which wrote through noodles' record writers.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "exon_b200_internal.h"

namespace exb {
int set_err(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
}  // namespace exb
using namespace exb;

namespace {

constexpr int64_t FMT_COUNTERS = 64;  // bytes reserved for the kernels' counters

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct Scratch {
    uint32_t* lens;
    int64_t* long_rows;
    unsigned long long* counters;
    void* scan_ws;
    int64_t scan_ws_bytes;
};
int64_t scratch_bytes(int64_t n_rows) {
    return FMT_COUNTERS + align256(4 * (n_rows + 1)) + align256(8 * (n_rows + 1)) + align256(exb_scan_workspace_bytes(4 * (n_rows + 1)));
}
int carve(void* d_scratch, int64_t bytes, int64_t n_rows, Scratch* s) {
    if (!d_scratch || bytes < scratch_bytes(n_rows)) return set_err(EXB_ERR_ARG, "format scratch too small: need %lld bytes", (long long)scratch_bytes(n_rows));
    uint8_t* p = reinterpret_cast<uint8_t*>(d_scratch);
    s->counters = reinterpret_cast<unsigned long long*>(p);
    p += FMT_COUNTERS;
    s->lens = reinterpret_cast<uint32_t*>(p);
    p += align256(4 * (n_rows + 1));
    s->long_rows = reinterpret_cast<int64_t*>(p);
    p += align256(8 * (n_rows + 1));
    s->scan_ws = p;
    s->scan_ws_bytes = align256(exb_scan_workspace_bytes(4 * (n_rows + 1)));
    return 0;
}

int format_columns(bool fasta, const exb_format_cols* c, int64_t n_rows, int line_width, int64_t* d_row_off, uint8_t* d_out, int64_t out_cap,
                   void* d_scratch, int64_t scratch_sz, void* stream) {
    const int n_cols = fasta ? 3 : 4;
    if (!c || n_rows < 0 || !d_row_off || out_cap < 0 || (out_cap > 0 && !d_out)) return set_err(EXB_ERR_ARG, "format: bad arguments");
    if (fasta && line_width < 1) return set_err(EXB_ERR_ARG, "format: line_width must be positive");
    for (int i = 0; i < n_cols; i++) {
        if (!c->d_off[i] || (n_rows > 0 && !c->d_data[i] && i != 1)) return set_err(EXB_ERR_ARG, "format: column %d is missing", i);
        if ((uintptr_t)c->d_data[i] & 3) return set_err(EXB_ERR_ARG, "format: column %d data must be 4-byte aligned", i);
    }
    cudaStream_t st = (cudaStream_t)stream;
    Scratch s;
    int rc = carve(d_scratch, scratch_sz, n_rows, &s);
    if (rc) return rc;
    FormatArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_cols; i++) {
        a.off[i] = c->d_off[i];
        a.data[i] = c->d_data[i];
    }
    a.desc_valid = c->d_desc_valid;
    a.n_rows = n_rows;
    a.line_width = line_width;
    a.lens = s.lens;
    a.row_off = d_row_off;
    a.out = d_out;
    a.out_cap = out_cap;
    a.long_rows = s.long_rows;
    a.counters = s.counters;
    cudaError_t e = format_len_launch(a, fasta, st);
    if (e != cudaSuccess) return cuda_fail(e, "format_len launch");
    rc = exb_exclusive_scan_u32(s.lens, n_rows, d_row_off, s.scan_ws, s.scan_ws_bytes, stream);
    if (rc) return rc;
    if (n_rows > 0) {
        e = format_rows_launch(a, fasta, st);
        if (e != cudaSuccess) return cuda_fail(e, "format_rows launch");
    }
    return 0;
}

// ---------------------------------------------------------------- output sinks
struct Sink {
    virtual ~Sink() {}
    virtual int write(const uint8_t* p, int64_t n) = 0;
    virtual int finish() = 0;
};
int write_all(int fd, const uint8_t* p, int64_t n) {
    while (n > 0) {
        const ssize_t w = ::write(fd, p, (size_t)(n > (1 << 30) ? (1 << 30) : n));
        if (w < 0) {
            if (errno == EINTR) continue;
            return set_err(EXB_ERR_IO, "write failed: %s", strerror(errno));
        }
        p += w;
        n -= w;
    }
    return 0;
}
struct PlainSink : Sink {
    int fd;
    explicit PlainSink(int f) : fd(f) {}
    ~PlainSink() override {
        if (fd >= 0) ::close(fd);
    }
    int write(const uint8_t* p, int64_t n) override { return write_all(fd, p, n); }
    int finish() override {
        const int rc = ::close(fd);
        fd = -1;
        return rc == 0 ? 0 : set_err(EXB_ERR_IO, "close failed: %s", strerror(errno));
    }
};
struct GzipSink : Sink {
    gzFile gz;
    explicit GzipSink(int fd) : gz(gzdopen(fd, "wb6")) {}
    ~GzipSink() override {
        if (gz) gzclose(gz);
    }
    int write(const uint8_t* p, int64_t n) override {
        while (n > 0) {
            const unsigned m = (unsigned)(n > (1 << 30) ? (1 << 30) : n);
            const int w = gzwrite(gz, p, m);
            if (w <= 0) return set_err(EXB_ERR_IO, "gzwrite failed");
            p += w;
            n -= w;
        }
        return 0;
    }
    int finish() override {
        const int rc = gzclose(gz);
        gz = nullptr;
        return rc == Z_OK ? 0 : set_err(EXB_ERR_IO, "gzclose failed (%d)", rc);
    }
};
// gzip output as BGZF: the image the device formatted is cut into members of 65280 text bytes and deflated on several host
// threads (exb_bgzf_compress_host, inflate.cu) -- a valid multi-member gzip file for every other reader, and one this
// library's own reader inflates on the device.  (EXON_B200_GZIP_WRITER=zlib keeps the single-stream sink above.)
struct BgzfSink : Sink {
    int fd;
    std::vector<uint8_t> carry, out;  // text that does not fill a member yet; compressed members of one write
    int threads;
    explicit BgzfSink(int f) : fd(f) {
        const unsigned hc = std::thread::hardware_concurrency();
        threads = (int)(hc ? (hc > 16 ? 16 : hc) : 4);
        if (const char* e = getenv("EXON_B200_WRITER_THREADS")) threads = atoi(e) > 0 ? atoi(e) : threads;
    }
    ~BgzfSink() override {
        if (fd >= 0) ::close(fd);
    }
    int emit(const uint8_t* p, int64_t n, bool eof) {
        out.resize((size_t)exb_bgzf_compress_bound(n));
        int64_t w = 0;
        if (exb_bgzf_compress_host(p, n, 6, threads, eof ? 1 : 0, out.data(), (int64_t)out.size(), &w) != 0) return EXB_ERR_IO;
        return write_all(fd, out.data(), w);
    }
    int write(const uint8_t* p, int64_t n) override {
        constexpr int64_t M = 65280;
        if (!carry.empty()) {  // complete the open member first
            const int64_t take = std::min<int64_t>(n, M - (int64_t)carry.size());
            carry.insert(carry.end(), p, p + take);
            p += take;
            n -= take;
            if ((int64_t)carry.size() < M) return 0;
            if (int rc = emit(carry.data(), M, false)) return rc;
            carry.clear();
        }
        const int64_t whole = n / M * M;
        if (whole)
            if (int rc = emit(p, whole, false)) return rc;
        carry.assign(p + whole, p + n);
        return 0;
    }
    int finish() override {
        int rc = emit(carry.data(), (int64_t)carry.size(), true);
        carry.clear();
        if (::close(fd) != 0 && rc == 0) rc = set_err(EXB_ERR_IO, "close failed: %s", strerror(errno));
        fd = -1;
        return rc;
    }
};
// zstd: the public streaming compression API of the system's libzstd.so.1 (same library the reader binds for input)
struct ZIn { const void* src; size_t size, pos; };
struct ZOut { void* dst; size_t size, pos; };
struct ZstdC {
    void* (*create)() = nullptr;
    size_t (*destroy)(void*) = nullptr;
    size_t (*compress2)(void*, ZOut*, ZIn*, int) = nullptr;
    size_t (*out_size)() = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    const char* (*error_name)(size_t) = nullptr;
    bool ok = false;
    static const ZstdC& get() {
        static const ZstdC l = [] {
            ZstdC z;
            void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
            if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
            if (!h) return z;
            z.create = reinterpret_cast<void* (*)()>(dlsym(h, "ZSTD_createCCtx"));
            z.destroy = reinterpret_cast<size_t (*)(void*)>(dlsym(h, "ZSTD_freeCCtx"));
            z.compress2 = reinterpret_cast<size_t (*)(void*, ZOut*, ZIn*, int)>(dlsym(h, "ZSTD_compressStream2"));
            z.out_size = reinterpret_cast<size_t (*)()>(dlsym(h, "ZSTD_CStreamOutSize"));
            z.is_error = reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError"));
            z.error_name = reinterpret_cast<const char* (*)(size_t)>(dlsym(h, "ZSTD_getErrorName"));
            z.ok = z.create && z.destroy && z.compress2 && z.out_size && z.is_error && z.error_name;
            return z;
        }();
        return l;
    }
};
struct ZstdSink : Sink {
    int fd;
    void* ctx;
    std::vector<uint8_t> out;
    explicit ZstdSink(int f) : fd(f), ctx(ZstdC::get().create()), out(ZstdC::get().out_size()) {}
    ~ZstdSink() override {
        if (ctx) ZstdC::get().destroy(ctx);
        if (fd >= 0) ::close(fd);
    }
    int pump(const uint8_t* p, int64_t n, int mode) {  // mode 0 = continue, 2 = end
        const ZstdC& z = ZstdC::get();
        ZIn in{p, (size_t)n, 0};
        for (;;) {
            ZOut o{out.data(), out.size(), 0};
            const size_t r = z.compress2(ctx, &o, &in, mode);
            if (z.is_error(r)) return set_err(EXB_ERR_IO, "zstd: %s", z.error_name(r));
            if (o.pos) {
                const int rc = write_all(fd, out.data(), (int64_t)o.pos);
                if (rc) return rc;
            }
            if (mode == 0 ? in.pos == in.size : r == 0) return 0;
        }
    }
    int write(const uint8_t* p, int64_t n) override { return pump(p, n, 0); }
    int finish() override {
        int rc = pump(nullptr, 0, 2);
        if (::close(fd) != 0 && rc == 0) rc = set_err(EXB_ERR_IO, "close failed: %s", strerror(errno));
        fd = -1;
        return rc;
    }
};

bool ends_with(const std::string& s, const char* suffix) {
    const size_t n = strlen(suffix);
    return s.size() >= n && strcasecmp(s.c_str() + s.size() - n, suffix) == 0;
}
// 0 = none, 1 = gzip, 2 = zstd, -1 = unknown
int parse_compression(const char* compression, const std::string& path) {
    if (!compression || !*compression || !strcasecmp(compression, "auto")) {
        if (ends_with(path, ".gz")) return 1;
        if (ends_with(path, ".zst")) return 2;
        return 0;
    }
    if (!strcasecmp(compression, "gzip") || !strcasecmp(compression, "gz")) return 1;
    if (!strcasecmp(compression, "zstd") || !strcasecmp(compression, "zst")) return 2;
    if (!strcasecmp(compression, "none") || !strcasecmp(compression, "uncompressed")) return 0;
    return -1;
}

template <class T>
struct Pinned {  // growable pinned host array + its device twin
    T* h = nullptr;
    T* d = nullptr;
    int64_t cap = 0;
    int reserve(int64_t n, int64_t keep) {  // keeps the first `keep` host elements
        if (n <= cap) return 0;
        int64_t want = cap ? cap : 1;
        while (want < n) want *= 2;
        T* nh = nullptr;
        T* nd = nullptr;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&nh), (size_t)want * sizeof(T), cudaHostAllocDefault);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc(writer staging)");
        e = cudaMalloc(reinterpret_cast<void**>(&nd), (size_t)want * sizeof(T) + 16);
        if (e != cudaSuccess) {
            cudaFreeHost(nh);
            return cuda_fail(e, "cudaMalloc(writer staging)");
        }
        if (keep > 0 && h) memcpy(nh, h, (size_t)keep * sizeof(T));
        release();
        h = nh;
        d = nd;
        cap = want;
        return 0;
    }
    void release() {
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr;
        cap = 0;
    }
};

}  // namespace

struct exb_writer {
    bool fasta = false;
    int n_cols = 4;
    int line_width = 80;
    int device = 0;
    Sink* sink = nullptr;
    cudaStream_t st = nullptr;
    // staged rows (host side pinned, device twins of the same capacity)
    Pinned<int64_t> off[4];
    Pinned<uint8_t> data[4];
    Pinned<uint8_t> valid;
    int64_t rows = 0;
    int64_t bytes[4] = {0, 0, 0, 0};
    // device scratch + image
    void* d_scratch = nullptr;
    int64_t scratch_cap = 0;
    int64_t* d_row_off = nullptr;
    int64_t row_off_cap = 0;
    Pinned<uint8_t> image;
    int64_t rows_written = 0, bytes_written = 0;
    int64_t flush_bytes = 64ll << 20, flush_rows = 4ll << 20;
    bool failed = false;

    ~exb_writer() {
        cudaSetDevice(device);
        for (int i = 0; i < 4; i++) {
            off[i].release();
            data[i].release();
        }
        valid.release();
        image.release();
        if (d_scratch) cudaFree(d_scratch);
        if (d_row_off) cudaFree(d_row_off);
        if (st) cudaStreamDestroy(st);
        delete sink;
    }

    int64_t image_bound() const {  // bytes the staged rows can occupy in the file
        int64_t b = 0;
        for (int i = 0; i < n_cols; i++) b += bytes[i];
        if (fasta) return b + 3 * rows + bytes[2] / line_width + rows;
        return b + 7 * rows;  // '@' ' ' '\n' '\n' '+' '\n' '\n'
    }

    int flush() {
        if (rows == 0) return 0;
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        const int64_t need_scratch = exb_format_scratch_bytes(rows);
        if (need_scratch > scratch_cap) {
            if (d_scratch) cudaFree(d_scratch);
            d_scratch = nullptr;
            scratch_cap = 0;
            if ((e = cudaMalloc(&d_scratch, (size_t)need_scratch)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(format scratch)");
            scratch_cap = need_scratch;
        }
        if (rows + 1 > row_off_cap) {
            if (d_row_off) cudaFree(d_row_off);
            d_row_off = nullptr;
            row_off_cap = 0;
            if ((e = cudaMalloc(reinterpret_cast<void**>(&d_row_off), (size_t)(rows + 1) * 8)) != cudaSuccess)
                return cuda_fail(e, "cudaMalloc(row offsets)");
            row_off_cap = rows + 1;
        }
        const int64_t bound = image_bound();
        int rc = image.reserve(bound > 0 ? bound : 1, 0);
        if (rc) return rc;
        exb_format_cols c;
        memset(&c, 0, sizeof(c));
        for (int i = 0; i < n_cols; i++) {
            if ((e = cudaMemcpyAsync(off[i].d, off[i].h, (size_t)(rows + 1) * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
                return cuda_fail(e, "H2D offsets");
            if (bytes[i] > 0 && (e = cudaMemcpyAsync(data[i].d, data[i].h, (size_t)bytes[i], cudaMemcpyHostToDevice, st)) != cudaSuccess)
                return cuda_fail(e, "H2D column");
            c.d_off[i] = off[i].d;
            c.d_data[i] = data[i].d;
        }
        if ((e = cudaMemcpyAsync(valid.d, valid.h, (size_t)rows, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D validity");
        c.d_desc_valid = valid.d;
        rc = fasta ? exb_fasta_format(&c, rows, line_width, d_row_off, image.d, bound, d_scratch, scratch_cap, st)
                   : exb_fastq_format(&c, rows, d_row_off, image.d, bound, d_scratch, scratch_cap, st);
        if (rc) return rc;
        int64_t out_bytes = 0;
        rc = exb_format_finish(d_row_off, rows, d_scratch, bound, &out_bytes, st);
        if (rc) return rc;
        if ((e = cudaMemcpyAsync(image.h, image.d, (size_t)out_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(e, "D2H image");
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "writer stream");
        rc = sink->write(image.h, out_bytes);
        if (rc) return rc;
        rows_written += rows;
        bytes_written += out_bytes;
        rows = 0;
        for (int i = 0; i < 4; i++) bytes[i] = 0;
        return 0;
    }
};

extern "C" {

int64_t exb_format_scratch_bytes(int64_t n_rows) { return scratch_bytes(n_rows < 0 ? 0 : n_rows); }

int exb_fastq_format(const exb_format_cols* cols, int64_t n_rows, int64_t* d_row_off, uint8_t* d_out, int64_t out_cap, void* d_scratch,
                     int64_t scratch_sz, void* stream) {
    return format_columns(false, cols, n_rows, 0, d_row_off, d_out, out_cap, d_scratch, scratch_sz, stream);
}
int exb_fasta_format(const exb_format_cols* cols, int64_t n_rows, int line_width, int64_t* d_row_off, uint8_t* d_out, int64_t out_cap,
                     void* d_scratch, int64_t scratch_sz, void* stream) {
    return format_columns(true, cols, n_rows, line_width, d_row_off, d_out, out_cap, d_scratch, scratch_sz, stream);
}
int exb_format_finish(const int64_t* d_row_off, int64_t n_rows, const void* d_scratch, int64_t out_cap, int64_t* out_bytes, void* stream) {
    if (!d_row_off || !d_scratch || !out_bytes || n_rows < 0) return set_err(EXB_ERR_ARG, "exb_format_finish: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long counters[4] = {0, 0, 0, 0};
    int64_t total = 0;
    cudaError_t e = cudaMemcpyAsync(counters, d_scratch, sizeof(counters), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_row_off + n_rows, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "exb_format_finish");
    *out_bytes = total;
    if (counters[1] != 0) return set_err(EXB_ERR_ARG, "record %llu would exceed 4 GiB in the file", (unsigned long long)~counters[1]);
    if (counters[2] != 0 || total > out_cap)
        return set_err(EXB_ERR_CAPACITY, "file image needs %lld bytes, the output buffer has %lld", (long long)total, (long long)out_cap);
    return 0;
}

int exb_writer_open(const char* path, const char* file_format, const char* compression, int force, int device, exb_writer** out) {
    if (!path || !*path || !file_format || !out) return set_err(EXB_ERR_ARG, "exb_writer_open: bad arguments");
    *out = nullptr;
    bool fasta;
    if (!strcasecmp(file_format, "fastq")) fasta = false;
    else if (!strcasecmp(file_format, "fasta")) fasta = true;
    else return set_err(EXB_ERR_ARG, "unknown file format '%s' (fasta | fastq)", file_format);
    const int comp = parse_compression(compression, path);
    if (comp < 0) return set_err(EXB_ERR_ARG, "unsupported compression '%s' for writing (gzip | zstd | none)", compression);
    if (comp == 2 && !ZstdC::get().ok) return set_err(EXB_ERR_IO, "libzstd.so.1 is not available on this host");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return set_err(EXB_ERR_CUDA, "no CUDA device: the exon_b200 writers format on the GPU");
    if (device < 0 || device >= n_dev) return set_err(EXB_ERR_ARG, "device %d does not exist (%d visible)", device, n_dev);
    struct stat sb;
    const bool exists = ::stat(path, &sb) == 0;
    if (exists && S_ISREG(sb.st_mode) && !force) return set_err(EXB_ERR_IO, "%s exists (FORCE true overwrites it)", path);
    const int fd = ::open(path, O_WRONLY | O_CREAT | (exists && !S_ISREG(sb.st_mode) ? 0 : O_TRUNC), 0644);
    if (fd < 0) return set_err(EXB_ERR_IO, "cannot create %s: %s", path, strerror(errno));
    exb_writer* w = new exb_writer();
    w->fasta = fasta;
    w->n_cols = fasta ? 3 : 4;
    w->device = device;
    const bool zlib_sink = getenv("EXON_B200_GZIP_WRITER") && !strcmp(getenv("EXON_B200_GZIP_WRITER"), "zlib");
    w->sink = comp == 1 ? (zlib_sink ? static_cast<Sink*>(new GzipSink(fd)) : static_cast<Sink*>(new BgzfSink(fd))) : comp == 2 ? static_cast<Sink*>(new ZstdSink(fd)) : static_cast<Sink*>(new PlainSink(fd));
    if (const char* e = getenv("EXON_B200_WRITER_FLUSH_BYTES")) w->flush_bytes = atoll(e) > 0 ? atoll(e) : w->flush_bytes;
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        delete w;
        return cuda_fail(ce, "exb_writer_open");
    }
    *out = w;
    return 0;
}

int exb_writer_set_line_width(exb_writer* w, int line_width) {
    if (!w || line_width < 1) return set_err(EXB_ERR_ARG, "exb_writer_set_line_width: bad arguments");
    if (w->rows) {
        const int rc = w->flush();
        if (rc) return rc;
    }
    w->line_width = line_width;
    return 0;
}

int exb_writer_append(exb_writer* w, int64_t n_rows, const int64_t* const* offsets, const uint8_t* const* data, const uint8_t* desc_valid) {
    if (!w || n_rows < 0 || !offsets || !data) return set_err(EXB_ERR_ARG, "exb_writer_append: bad arguments");
    if (w->failed) return set_err(EXB_ERR_IO, "the writer failed earlier");
    if (n_rows == 0) return 0;
    for (int i = 0; i < w->n_cols; i++)
        if (!offsets[i] || offsets[i][n_rows] < offsets[i][0]) return set_err(EXB_ERR_ARG, "exb_writer_append: column %d has no valid offsets", i);
    cudaSetDevice(w->device);
    int64_t add = 0;
    for (int i = 0; i < w->n_cols; i++) add += offsets[i][n_rows] - offsets[i][0];
    int64_t staged = 0;
    for (int i = 0; i < w->n_cols; i++) staged += w->bytes[i];
    if (w->rows > 0 && (staged + add > w->flush_bytes || w->rows + n_rows > w->flush_rows)) {
        const int rc = w->flush();
        if (rc) {
            w->failed = true;
            return rc;
        }
    }
    int rc = w->valid.reserve(w->rows + n_rows, w->rows);
    for (int i = 0; i < w->n_cols && !rc; i++) {
        const int64_t nb = offsets[i][n_rows] - offsets[i][0];
        rc = w->off[i].reserve(w->rows + n_rows + 1, w->rows + 1);
        if (!rc) rc = w->data[i].reserve(w->bytes[i] + nb + 4, w->bytes[i]);
    }
    if (rc) {
        w->failed = true;
        return rc;
    }
    for (int i = 0; i < w->n_cols; i++) {
        const int64_t b0 = offsets[i][0], nb = offsets[i][n_rows] - b0;
        int64_t* o = w->off[i].h + w->rows;
        const int64_t shift = w->bytes[i] - b0;
        for (int64_t r = 0; r <= n_rows; r++) o[r] = offsets[i][r] + shift;
        if (nb > 0) memcpy(w->data[i].h + w->bytes[i], data[i] + b0, (size_t)nb);
        w->bytes[i] += nb;
    }
    if (desc_valid) memcpy(w->valid.h + w->rows, desc_valid, (size_t)n_rows);
    else memset(w->valid.h + w->rows, 1, (size_t)n_rows);
    w->rows += n_rows;
    return 0;
}

int exb_writer_close(exb_writer* w, int64_t* rows_written, int64_t* bytes_written) {
    if (!w) return set_err(EXB_ERR_ARG, "exb_writer_close: no writer");
    int rc = 0;
    if (!w->failed) rc = w->flush();
    const int rc2 = w->sink->finish();
    if (!rc) rc = rc2;
    if (rows_written) *rows_written = w->rows_written;
    if (bytes_written) *bytes_written = w->bytes_written;
    delete w;
    return rc;
}

}  // extern "C"
