// reader.cu -- the reference-FFI layer: new_reader / replacement_scan
// (exon/include/rust.hpp:41-48; rust/src/arrow_reader.rs:38-197).
//
// new_reader opens a FASTA/FASTQ file (or a directory of them) and returns an
// Arrow C stream of record batches, exactly what the reference's C++ glue
// consumes (arrow_table_function/module.cpp:95-112,235-252; duckdb
// arrow.cpp:248-268).  Behind the stream, file bytes go host (pinned) -> HBM
// in chunks; each chunk is scanned, filtered, split into fields and compacted
// into column buffers by the kernels of this library, and only the resulting
// columns travel back.  A record that straddles a chunk edge is re-read with
// the next chunk (the scan reports where the last complete record ended).
#include <ctype.h>
#include <cuda_runtime.h>
#include <dirent.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <immintrin.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "exon_b200_internal.h"

namespace exb {
int set_err(int code, const char* fmt, ...);
cudaError_t bgzf_inflate_launch(const uint8_t* d_in, const exb_bgzf_block* d_blocks, int64_t n_blocks, uint8_t* d_out, unsigned int* d_state, int check_crc,
                                int64_t block_base, cudaStream_t st);
cudaError_t str_pred_launch(const uint8_t*, const int64_t*, const uint32_t*, const uint8_t*, int64_t, int, const uint8_t*, int, uint8_t*,
                            cudaStream_t);
cudaError_t pass_combine_launch(uint8_t*, const uint8_t*, int64_t, int, cudaStream_t);
cudaError_t fasta_num_pred_launch(const int64_t*, const int64_t*, int64_t, int, int, double, uint8_t*, cudaStream_t);
cudaError_t fasta_seq_ranges_launch(const int64_t*, const int64_t*, int64_t, int64_t*, uint32_t*, cudaStream_t);
cudaError_t take_u32_launch(const uint32_t*, const int64_t*, int64_t, uint32_t*, cudaStream_t);
cudaError_t take_i64_launch(const int64_t*, const int64_t*, int64_t, int64_t*, cudaStream_t);
cudaError_t take_u8_launch(const uint8_t*, const int64_t*, int64_t, uint8_t*, cudaStream_t);
// reader_ops.cu
cudaError_t string_t_launch(const int64_t*, const uint8_t*, uint64_t, int64_t, void*, cudaStream_t);
cudaError_t string_t_borrow_launch(const int64_t*, const uint32_t*, const uint8_t*, uint64_t, int64_t, void*, cudaStream_t);
cudaError_t valid_bits_launch(const uint8_t*, int64_t, uint64_t*, uint64_t*, cudaStream_t);
cudaError_t list_entries_launch(const int64_t*, int64_t, int64_t, void*, int64_t*, cudaStream_t);
cudaError_t gc_sel_launch(const uint32_t*, const uint32_t*, const int64_t*, const int64_t*, const int64_t*, int64_t, float*, cudaStream_t);
cudaError_t mean_quality_launch(const uint32_t*, const int32_t*, const int64_t*, int64_t, double*, uint8_t*, cudaStream_t);
cudaError_t lens_i64_launch(const uint32_t*, int64_t, int64_t*, cudaStream_t);
cudaError_t fastq_chunk_info_launch(const void*, const uint32_t*, int64_t*, cudaStream_t);
cudaError_t copy_words_launch(void*, const void*, int, int64_t, cudaStream_t);
cudaError_t gather_ranges_map_launch(const uint8_t*, const int64_t*, const int64_t*, int64_t, int64_t, uint8_t*, int, unsigned long long*, cudaStream_t);
}  // namespace exb
using namespace exb;

static char* dup_cstr(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
static std::string lower(std::string s) {
    for (auto& c : s) c = (char)tolower((unsigned char)c);
    return s;
}
static bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

// ExonFileType::from_str (exon 0.2.6, not in the reference tree): case-insensitive
// "fasta" | "fa" | "fna" -> FASTA, "fastq" | "fq" -> FASTQ.  Pinned by the
// reference tests for ".fasta" and ".fastq" only (test_fasta_scan.test:28-43,
// test_fastq_scan.test:43-59).
static int file_type_from_str(const std::string& s) {
    std::string l = lower(s);
    if (l == "fasta" || l == "fa" || l == "fna") return 1;
    if (l == "fastq" || l == "fq") return 2;
    return 0;
}
// FileCompressionType::from_str (datafusion 28): GZIP|GZ, ZSTD|ZST, BZIP2|BZ2, XZ (case-insensitive)
static int compression_from_str(const std::string& s) {
    std::string l = lower(s);
    if (l == "gzip" || l == "gz") return 1;
    if (l == "zstd" || l == "zst") return 2;
    if (l == "bzip2" || l == "bz2") return 3;
    if (l == "xz") return 4;
    return 0;
}

// ------------------------------------------------------------------ zstd input
// The streaming decoder of the system's libzstd (zstd >= 1.0 ABI), bound at run time: the image ships libzstd.so.1 but no
// header, and the reference's own copy lives in crates that are not on disk.  The declarations below restate the
// public, stable streaming API (zstd.h "Streaming decompression - HowTo").  Replaces async-compression's ZstdDecoder
// behind datafusion's FileCompressionType::ZSTD (arrow_reader.rs:60-91).
struct ZstdIn { const void* src; size_t size, pos; };
struct ZstdOut { void* dst; size_t size, pos; };
struct ZstdLib {
    void* (*create)() = nullptr;
    size_t (*destroy)(void*) = nullptr;
    size_t (*init)(void*) = nullptr;
    size_t (*decompress)(void*, ZstdOut*, ZstdIn*) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    const char* (*error_name)(size_t) = nullptr;
    bool ok = false;
    static const ZstdLib& get() {
        static ZstdLib L = [] {
            ZstdLib l;
            void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
            if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
            if (!h) return l;
            l.create = reinterpret_cast<void* (*)()>(dlsym(h, "ZSTD_createDStream"));
            l.destroy = reinterpret_cast<size_t (*)(void*)>(dlsym(h, "ZSTD_freeDStream"));
            l.init = reinterpret_cast<size_t (*)(void*)>(dlsym(h, "ZSTD_initDStream"));
            l.decompress = reinterpret_cast<size_t (*)(void*, ZstdOut*, ZstdIn*)>(dlsym(h, "ZSTD_decompressStream"));
            l.is_error = reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError"));
            l.error_name = reinterpret_cast<const char* (*)(size_t)>(dlsym(h, "ZSTD_getErrorName"));
            l.ok = l.create && l.destroy && l.init && l.decompress && l.is_error && l.error_name;
            return l;
        }();
        return L;
    }
};

// bzip2 and xz the same way (FileCompressionType::BZIP2 / XZ of datafusion 28: reachable through the `compression`
// option only, arrow_reader.rs:60-91 infers nothing but gz / zst from the file name).  Public, stable C APIs of
// libbz2 1.0 (bzlib.h) and liblzma 5 (lzma/base.h); LZMA_STREAM_INIT is all zeros and the struct is over-allocated.
struct Bz2Stream {
    char* next_in; unsigned int avail_in, total_in_lo32, total_in_hi32;
    char* next_out; unsigned int avail_out, total_out_lo32, total_out_hi32;
    void* state; void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};
struct XzStream {
    const uint8_t* next_in; size_t avail_in; uint64_t total_in;
    uint8_t* next_out; size_t avail_out; uint64_t total_out;
    const void* allocator; void* internal; void* reserved_ptr[4]; uint64_t reserved_int[2]; size_t reserved_size[2]; int reserved_enum[2];
    uint8_t slack[64];
};
struct Bz2Lib {
    int (*init)(Bz2Stream*, int, int) = nullptr;
    int (*run)(Bz2Stream*) = nullptr;
    int (*end)(Bz2Stream*) = nullptr;
    bool ok = false;
    static const Bz2Lib& get() {
        static Bz2Lib L = [] {
            Bz2Lib l;
            void* h = dlopen("libbz2.so.1.0", RTLD_NOW | RTLD_LOCAL);
            if (!h) h = dlopen("libbz2.so.1", RTLD_NOW | RTLD_LOCAL);
            if (!h) return l;
            l.init = reinterpret_cast<int (*)(Bz2Stream*, int, int)>(dlsym(h, "BZ2_bzDecompressInit"));
            l.run = reinterpret_cast<int (*)(Bz2Stream*)>(dlsym(h, "BZ2_bzDecompress"));
            l.end = reinterpret_cast<int (*)(Bz2Stream*)>(dlsym(h, "BZ2_bzDecompressEnd"));
            l.ok = l.init && l.run && l.end;
            return l;
        }();
        return L;
    }
};
struct XzLib {
    int (*init)(XzStream*, uint64_t, uint32_t) = nullptr;
    int (*code)(XzStream*, int) = nullptr;
    void (*end)(XzStream*) = nullptr;
    bool ok = false;
    static const XzLib& get() {
        static XzLib L = [] {
            XzLib l;
            void* h = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
            if (!h) return l;
            l.init = reinterpret_cast<int (*)(XzStream*, uint64_t, uint32_t)>(dlsym(h, "lzma_stream_decoder"));
            l.code = reinterpret_cast<int (*)(XzStream*, int)>(dlsym(h, "lzma_code"));
            l.end = reinterpret_cast<void (*)(XzStream*)>(dlsym(h, "lzma_end"));
            l.ok = l.init && l.code && l.end;
            return l;
        }();
        return L;
    }
};

// One streaming decoder behind the IO thread: consume in[ipos, isize), produce out[opos, ocap).  in_stream() is true while
// a compressed stream is open (end of file then = truncated input); concatenated streams / frames simply continue.
struct Decoder {
    virtual ~Decoder() {}
    virtual bool step(const uint8_t* in, size_t isize, size_t& ipos, uint8_t* out, size_t ocap, size_t& opos, std::string& err) = 0;
    virtual bool in_stream() const = 0;
    virtual const char* name() const = 0;
};
struct ZstdDecoder : Decoder {
    void* ds = nullptr;
    bool open_ = false;
    bool init(std::string& err) {
        const ZstdLib& Z = ZstdLib::get();
        if (!Z.ok) { err = "zstd input needs the system's libzstd.so.1, which could not be loaded"; return false; }
        if (!(ds = Z.create()) || Z.is_error(Z.init(ds))) { err = "could not create a zstd decoder"; return false; }
        return true;
    }
    ~ZstdDecoder() override { if (ds) ZstdLib::get().destroy(ds); }
    bool step(const uint8_t* in, size_t isize, size_t& ipos, uint8_t* out, size_t ocap, size_t& opos, std::string& err) override {
        const ZstdLib& Z = ZstdLib::get();
        ZstdIn zin{in, isize, ipos};
        ZstdOut zout{out, ocap, opos};
        const size_t r = Z.decompress(ds, &zout, &zin);  // 0 = a frame ended; the next call starts the next frame
        if (Z.is_error(r)) { err = std::string("zstd: ") + Z.error_name(r); return false; }
        open_ = r != 0;
        ipos = zin.pos;
        opos = zout.pos;
        return true;
    }
    bool in_stream() const override { return open_; }
    const char* name() const override { return "zstd"; }
};
struct Bz2Decoder : Decoder {
    Bz2Stream st;
    bool live = false, open_ = false;
    Bz2Decoder() { memset(&st, 0, sizeof(st)); }
    bool init(std::string& err) {
        if (!Bz2Lib::get().ok) { err = "bzip2 input needs the system's libbz2.so.1.0, which could not be loaded"; return false; }
        return true;
    }
    ~Bz2Decoder() override { if (live) Bz2Lib::get().end(&st); }
    bool step(const uint8_t* in, size_t isize, size_t& ipos, uint8_t* out, size_t ocap, size_t& opos, std::string& err) override {
        const Bz2Lib& B = Bz2Lib::get();
        if (!live) {  // first stream, or the one after a stream that ended (bzip2 files may be concatenated)
            memset(&st, 0, sizeof(st));
            if (B.init(&st, 0, 0) != 0) { err = "could not create a bzip2 decoder"; return false; }
            live = true;
        }
        const size_t ni = std::min<size_t>(isize - ipos, 1u << 30), no = std::min<size_t>(ocap - opos, 1u << 30);
        st.next_in = reinterpret_cast<char*>(const_cast<uint8_t*>(in + ipos));
        st.avail_in = (unsigned)ni;
        st.next_out = reinterpret_cast<char*>(out + opos);
        st.avail_out = (unsigned)no;
        const int r = B.run(&st);
        if (r != 0 && r != 4) { err = "bzip2: corrupt input (code " + std::to_string(r) + ")"; return false; }  // BZ_OK / BZ_STREAM_END
        if (ni - st.avail_in > 0 || no - st.avail_out > 0) open_ = true;
        ipos += ni - st.avail_in;
        opos += no - st.avail_out;
        if (r == 4) {
            B.end(&st);
            live = false;
            open_ = false;
        }
        return true;
    }
    bool in_stream() const override { return open_; }
    const char* name() const override { return "bzip2"; }
};
struct XzDecoder : Decoder {
    XzStream st;
    bool live = false, open_ = false;
    XzDecoder() { memset(&st, 0, sizeof(st)); }
    bool init(std::string& err) {
        if (!XzLib::get().ok) { err = "xz input needs the system's liblzma.so.5, which could not be loaded"; return false; }
        return true;
    }
    ~XzDecoder() override { if (live) XzLib::get().end(&st); }
    bool step(const uint8_t* in, size_t isize, size_t& ipos, uint8_t* out, size_t ocap, size_t& opos, std::string& err) override {
        const XzLib& X = XzLib::get();
        if (!live) {
            while (ipos < isize && in[ipos] == 0) ipos++;  // stream padding between concatenated .xz streams
            if (ipos == isize) return true;
            memset(&st, 0, sizeof(st));
            if (X.init(&st, ~0ull, 0) != 0) { err = "could not create an xz decoder"; return false; }
            live = true;
        }
        st.next_in = in + ipos;
        st.avail_in = isize - ipos;
        st.next_out = out + opos;
        st.avail_out = ocap - opos;
        const int r = X.code(&st, 0 /* LZMA_RUN */);
        const size_t used = (isize - ipos) - st.avail_in, made = (ocap - opos) - st.avail_out;
        // LZMA_OK 0, STREAM_END 1, NO_CHECK 2, UNSUPPORTED_CHECK 3, GET_CHECK 4 carry on; BUF_ERROR 10 = "nothing to do with what
        // it was given", harmless when it only means that more input is needed
        if (r > 4 && !(r == 10 && used == 0 && made == 0)) { err = "xz: corrupt input (code " + std::to_string(r) + ")"; return false; }
        if (used || made) open_ = true;
        ipos += used;
        opos += made;
        if (r == 1) {
            X.end(&st);
            live = false;
            open_ = false;
        }
        return true;
    }
    bool in_stream() const override { return open_; }
    const char* name() const override { return "xz"; }
};

// ------------------------------------------------------------------ filter expressions
namespace {

enum NodeKind { N_AND, N_OR, N_STR, N_NUM };
struct Node {
    NodeKind kind;
    int lhs = -1, rhs = -1;  // children (AND / OR)
    int col = -1;            // N_STR: column index
    int op = 0;              // EXB_OP_*, 6 IS NULL, 7 IS NOT NULL
    std::string cst;         // N_STR constant
    int field = 0;           // N_NUM: EXB_P_*
    double value = 0;
};

struct Parser {
    const std::string& s;
    size_t i = 0;
    std::vector<Node>& nodes;
    const std::vector<std::string>& cols;
    std::string err;
    Parser(const std::string& s_, std::vector<Node>& n, const std::vector<std::string>& c) : s(s_), nodes(n), cols(c) {}

    void ws() {
        while (i < s.size() && isspace((unsigned char)s[i])) i++;
    }
    bool kw(const char* k) {  // case-insensitive keyword followed by a non-identifier char
        ws();
        size_t n = strlen(k);
        if (i + n > s.size()) return false;
        for (size_t j = 0; j < n; j++)
            if (toupper((unsigned char)s[i + j]) != k[j]) return false;
        if (i + n < s.size() && (isalnum((unsigned char)s[i + n]) || s[i + n] == '_')) return false;
        i += n;
        return true;
    }
    std::string ident() {
        ws();
        size_t b = i;
        if (i < s.size() && s[i] == '"') {  // quoted identifier
            i++;
            b = i;
            while (i < s.size() && s[i] != '"') i++;
            std::string r = s.substr(b, i - b);
            if (i < s.size()) i++;
            return r;
        }
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_')) i++;
        return s.substr(b, i - b);
    }
    int parse_op() {
        ws();
        if (s.compare(i, 2, ">=") == 0) { i += 2; return EXB_OP_GE; }
        if (s.compare(i, 2, "<=") == 0) { i += 2; return EXB_OP_LE; }
        if (s.compare(i, 2, "!=") == 0) { i += 2; return EXB_OP_NE; }
        if (s.compare(i, 2, "<>") == 0) { i += 2; return EXB_OP_NE; }
        if (s.compare(i, 1, ">") == 0) { i += 1; return EXB_OP_GT; }
        if (s.compare(i, 1, "<") == 0) { i += 1; return EXB_OP_LT; }
        if (s.compare(i, 1, "=") == 0) { i += 1; return EXB_OP_EQ; }
        return -1;
    }
    int col_index(const std::string& name) {
        std::string l = lower(name);
        for (size_t c = 0; c < cols.size(); c++)
            if (cols[c] == l) return (int)c;
        return -1;
    }
    int leaf() {
        ws();
        if (i < s.size() && s[i] == '(') {
            i++;
            int e = parse_or();
            ws();
            if (e < 0) return -1;
            if (i >= s.size() || s[i] != ')') { err = "expected ')'"; return -1; }
            i++;
            return e;
        }
        std::string id = ident();
        if (id.empty()) { err = "expected a column or function name at offset " + std::to_string(i); return -1; }
        ws();
        Node n;
        if (i < s.size() && s[i] == '(') {  // function predicate
            std::string fn = lower(id);
            i++;
            std::string arg = lower(ident());
            ws();
            if (i >= s.size() || s[i] != ')') { err = "expected ')' after function argument"; return -1; }
            i++;
            n.kind = N_NUM;
            if (fn == "mean_quality" && arg == "quality_scores") n.field = EXB_P_MEAN_QUALITY;
            else if (fn == "gc_content" && arg == "sequence") n.field = EXB_P_GC_CONTENT;
            else if (fn == "length" && arg == "sequence") n.field = EXB_P_SEQ_LEN;
            else if (fn == "length" && arg == "quality_scores") n.field = EXB_P_QUAL_LEN;
            else { err = "unsupported function predicate " + fn + "(" + arg + ")"; return -1; }
            n.op = parse_op();
            if (n.op < 0) { err = "expected a comparison operator"; return -1; }
            ws();
            char* end = nullptr;
            n.value = strtod(s.c_str() + i, &end);
            if (end == s.c_str() + i) { err = "expected a number"; return -1; }
            i = end - s.c_str();
            nodes.push_back(n);
            return (int)nodes.size() - 1;
        }
        n.kind = N_STR;
        n.col = col_index(id);
        if (n.col < 0) { err = "unknown column " + id; return -1; }
        if (kw("IS")) {
            bool neg = kw("NOT");
            if (!kw("NULL")) { err = "expected NULL"; return -1; }
            n.op = neg ? 7 : 6;
            nodes.push_back(n);
            return (int)nodes.size() - 1;
        }
        n.op = parse_op();
        if (n.op < 0) { err = "expected a comparison operator after " + id; return -1; }
        ws();
        if (i >= s.size() || s[i] != '\'') { err = "expected a string constant"; return -1; }
        i++;
        while (i < s.size()) {  // Value::ToSQLString doubles embedded quotes
            if (s[i] == '\'') {
                if (i + 1 < s.size() && s[i + 1] == '\'') { n.cst.push_back('\''); i += 2; continue; }
                break;
            }
            n.cst.push_back(s[i++]);
        }
        if (i >= s.size()) { err = "unterminated string constant"; return -1; }
        i++;
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }
    int parse_and() {
        int l = leaf();
        while (l >= 0 && kw("AND")) {
            int r = leaf();
            if (r < 0) return -1;
            Node n;
            n.kind = N_AND;
            n.lhs = l;
            n.rhs = r;
            nodes.push_back(n);
            l = (int)nodes.size() - 1;
        }
        return l;
    }
    int parse_or() {
        int l = parse_and();
        while (l >= 0 && kw("OR")) {
            int r = parse_and();
            if (r < 0) return -1;
            Node n;
            n.kind = N_OR;
            n.lhs = l;
            n.rhs = r;
            nodes.push_back(n);
            l = (int)nodes.size() - 1;
        }
        return l;
    }
};


// ------------------------------------------------------------------ buffers
// Device buffers are recycled through a process-wide pool, one per device: cudaMalloc / cudaFree are synchronising driver
// calls that took anything from 0.1 ms to 1 s on the GPU box (VM), once per buffer per query -- a COUNT(*) whose pipeline
// ran for 150 ms was observed to take 0.4-1.4 s because of them (profiles/round2_e2e_stalls.txt).
struct DevPool {
    struct Slab {
        void* p;
        int64_t cap;
    };
    std::mutex mu;
    std::vector<Slab> idle[16];
    int64_t idle_bytes[16] = {0};
    static constexpr int64_t kMaxIdle = 3ll << 30;  // per device
    static DevPool& get() {
        static DevPool* p = new DevPool();  // never destroyed: freeing at exit would race the CUDA runtime's teardown
        return *p;
    }
    static int device() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) d = 0;
        return d < 0 || d >= 16 ? 0 : d;
    }
    void* take(int64_t want, int64_t* cap_out) {
        const int d = device();
        {
            std::lock_guard<std::mutex> lk(mu);
            int best = -1;
            for (int i = 0; i < (int)idle[d].size(); i++)
                if (idle[d][i].cap >= want && idle[d][i].cap <= 4 * want + (1 << 20) && (best < 0 || idle[d][i].cap < idle[d][best].cap)) best = i;
            if (best >= 0) {
                Slab sl = idle[d][best];
                idle[d].erase(idle[d].begin() + best);
                idle_bytes[d] -= sl.cap;
                *cap_out = sl.cap;
                return sl.p;
            }
        }
        void* p = nullptr;
        if (cudaMalloc(&p, (size_t)want) != cudaSuccess) {
            cudaGetLastError();
            trim(d, 0);  // give the idle slabs back and try once more
            if (cudaMalloc(&p, (size_t)want) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        *cap_out = want;
        return p;
    }
    void give(void* p, int64_t cap) {
        if (!p) return;
        const int d = device();
        {
            std::lock_guard<std::mutex> lk(mu);
            if (idle_bytes[d] + cap <= kMaxIdle) {
                idle[d].push_back(Slab{p, cap});
                idle_bytes[d] += cap;
                return;
            }
        }
        cudaFree(p);
    }
    void trim(int d, int64_t keep) {
        std::vector<Slab> drop;
        {
            std::lock_guard<std::mutex> lk(mu);
            while (idle_bytes[d] > keep && !idle[d].empty()) {
                drop.push_back(idle[d].back());
                idle_bytes[d] -= idle[d].back().cap;
                idle[d].pop_back();
            }
        }
        for (Slab& sl : drop) cudaFree(sl.p);
    }
};
struct DBuf {  // growable device buffer (the device thread of a reader owns it: its current device is the buffer's device)
    void* p = nullptr;
    int64_t cap = 0;
    ~DBuf() { release(); }
    void release() {
        if (p) DevPool::get().give(p, cap);
        p = nullptr;
        cap = 0;
    }
    bool need(int64_t n) {
        if (n <= cap) return true;
        if (p) cudaDeviceSynchronize();  // growing (rare): work in flight may still use the old slab, and cudaFree no longer waits for it
        release();
        int64_t want = std::max<int64_t>(n + n / 4 + 256, 4096);
        want = (want + 65535) & ~(int64_t)65535;
        p = DevPool::get().take(want, &cap);
        if (!p) {
            cap = 0;
            return false;
        }
        return true;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
struct HBuf {  // growable pinned host buffer (portable: every device of a multi-GPU scan may copy from / to it)
    void* p = nullptr;
    int64_t cap = 0;
    ~HBuf() { if (p) cudaFreeHost(p); }
    bool need(int64_t n) {
        if (n <= cap) return true;
        int64_t want = std::max<int64_t>(n + n / 4 + 256, 4096);
        void* q = nullptr;
        if (cudaHostAlloc(&q, (size_t)want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
        if (p) cudaFreeHost(p);
        p = q;
        cap = want;
        return true;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

// Pinned host buffers are expensive to create (cudaHostAlloc pins every page) and cheap to reuse: input blocks and
// result buffers cycle through this pool.  Shared by every reader of the process and by every ChunkResult still held
// by the host, so a batch may outlive the reader that produced it.
struct PinnedPool {
    std::mutex mu;
    std::vector<HBuf*> idle;
    ~PinnedPool() {
        for (HBuf* b : idle) delete b;
    }
    HBuf* get(int64_t bytes) {
        HBuf* b = nullptr;
        {
            std::lock_guard<std::mutex> lk(mu);
            int best = -1, largest = -1;
            for (int i = 0; i < (int)idle.size(); i++) {
                if (idle[i]->cap >= bytes && (best < 0 || idle[i]->cap < idle[best]->cap)) best = i;
                if (largest < 0 || idle[i]->cap > idle[largest]->cap) largest = i;
            }
            // a buffer more than 4x too large stays for a request of its own size class (input blocks vs. small metadata)
            if (best >= 0 && idle[best]->cap > 4 * bytes + (1 << 20)) {
                int small = -1;
                for (int i = 0; i < (int)idle.size(); i++)
                    if (idle[i]->cap < bytes && (small < 0 || idle[i]->cap > idle[small]->cap)) small = i;
                if (small >= 0) best = small;  // grow a small one instead
            }
            if (best < 0) best = largest;  // none fits: grow the largest one instead of leaving it idle forever
            if (best >= 0) {
                b = idle[best];
                idle.erase(idle.begin() + best);
            }
        }
        if (!b) b = new HBuf();
        if (!b->need(bytes)) {
            delete b;
            return nullptr;
        }
        return b;
    }
    void put(HBuf* b) {
        if (!b) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            int64_t held = b->cap;
            for (HBuf* x : idle) held += x->cap;
            if (held <= max_idle_bytes) {
                idle.push_back(b);
                return;
            }
        }
        delete b;  // over the cap: unpin it
    }
    // One pool per process: pinning a 64 MiB block costs ~10 ms, so a query that opens a reader right after another
    // one (DuckDB: bind opens one for the schema, init_global the real one) starts with warm buffers.
    int64_t max_idle_bytes = 4ll << 30;
    static std::shared_ptr<PinnedPool> shared() {
        static std::mutex m;
        static std::shared_ptr<PinnedPool> keep;  // keeps the buffers pinned for the life of the process
        std::lock_guard<std::mutex> lk(m);
        if (!keep) keep = std::make_shared<PinnedPool>();
        return keep;
    }
};

// ------------------------------------------------------------------ host copy workers
// File bytes reach the pinned block through the page cache: mmap + memcpy by a pool of worker threads.  Measured on the
// B200 box (tools/iobench.cu, profiles/round2_iobench.txt): pread into pinned memory tops out at 43 GB/s with 16
// threads (one copy_to_user per call), memcpy from a mapping of the same tmpfs file runs at 88 GB/s -- above the
// 55 GB/s of the PCIe link, so the H2D copy, not the host, bounds the scan.  One pool per process, shared by the
// readers of a multi-GPU scan.
struct IoPool {
    struct Job {
        std::atomic<int64_t> left{0};
        std::mutex mu;
        std::condition_variable cv;
    };
    struct Task {
        uint8_t* dst;
        const uint8_t* src;
        size_t n;
        Job* job;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Task> q;
    int nthreads = 1;
    IoPool() {
        int hw = (int)std::thread::hardware_concurrency();
        // three quarters of the cores (at most 12): 8 threads already copy faster than the PCIe link takes the bytes on the
        // 16-core box, 12 are steadier, and the device thread and the host's consumers need cores to keep the GPU fed
        nthreads = hw > 0 ? std::max(2, std::min(hw * 3 / 4, 12)) : 4;
        // one process per GPU on one box (torchrun sets LOCAL_WORLD_SIZE): the processes share the cores
        if (const char* lws = getenv("LOCAL_WORLD_SIZE"))
            if (atoi(lws) > 1 && hw > 0) nthreads = std::max(2, std::min(nthreads, hw * 3 / 4 / atoi(lws)));
        if (const char* e = getenv("EXON_B200_IO_THREADS"))
            if (atoi(e) > 0) nthreads = std::min(atoi(e), 64);
        for (int i = 0; i < nthreads - 1; i++) std::thread([this] { work(); }).detach();  // the caller is the last worker
    }
    static IoPool& get() {
        static IoPool* p = new IoPool();  // never destroyed: detached workers may outlive static destructors
        return *p;
    }
    // 64-byte streaming stores: the pinned block is written once and next read by the DMA engine, so there is nothing to
    // keep in the CPU caches.  Measured pipelined with the H2D copy (tools/iobench2.cu, profiles/round2_iobench2.txt):
    // 51.6 GB/s with 8 threads against 47.5 GB/s for memcpy and 44.3 GB/s for cached 64-byte stores.
    __attribute__((target("avx512f"))) static void copy_stream512(uint8_t* d, const uint8_t* s, size_t n) {
        size_t head = (64 - ((uintptr_t)d & 63)) & 63;
        if (head > n) head = n;
        if (head) memcpy(d, s, head);
        size_t i = head;
        for (; i + 256 <= n; i += 256) {
            const __m512i a = _mm512_loadu_si512((const void*)(s + i)), b = _mm512_loadu_si512((const void*)(s + i + 64)),
                          c = _mm512_loadu_si512((const void*)(s + i + 128)), e = _mm512_loadu_si512((const void*)(s + i + 192));
            _mm512_stream_si512((__m512i*)(d + i), a);
            _mm512_stream_si512((__m512i*)(d + i + 64), b);
            _mm512_stream_si512((__m512i*)(d + i + 128), c);
            _mm512_stream_si512((__m512i*)(d + i + 192), e);
        }
        for (; i + 64 <= n; i += 64) _mm512_stream_si512((__m512i*)(d + i), _mm512_loadu_si512((const void*)(s + i)));
        _mm_sfence();
        if (i < n) memcpy(d + i, s + i, n - i);
    }
    static bool have_avx512() {
        static const bool yes = __builtin_cpu_supports("avx512f") && !getenv("EXON_B200_NO_AVX512");
        return yes;
    }
    static void run(const Task& t) {
        if (have_avx512()) copy_stream512(t.dst, t.src, t.n);
        else memcpy(t.dst, t.src, t.n);
        if (t.job->left.fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(t.job->mu);
            t.job->cv.notify_all();
        }
    }
    void work() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !q.empty(); });
                t = q.front();
                q.pop_front();
            }
            run(t);
        }
    }
    void copy(uint8_t* dst, const uint8_t* src, int64_t n) {
        if (n <= 0) return;
        if (n < (4 << 20) || nthreads == 1) {
            Job one;
            one.left.store(1);
            run(Task{dst, src, (size_t)n, &one});
            return;
        }
        int64_t slice = (n + nthreads - 1) / nthreads;
        slice = std::max<int64_t>((slice + 4095) & ~(int64_t)4095, 1 << 20);
        const int64_t parts = (n + slice - 1) / slice;
        Job job;
        job.left.store(parts);
        {
            std::lock_guard<std::mutex> lk(mu);
            for (int64_t k = 0; k < parts; k++) q.push_back(Task{dst + k * slice, src + k * slice, (size_t)std::min(slice, n - k * slice), &job});
        }
        cv.notify_all();
        for (;;) {  // help until the queue is empty, then wait for the slices other workers hold
            Task t;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (q.empty()) break;
                t = q.front();
                q.pop_front();
            }
            run(t);
        }
        std::unique_lock<std::mutex> lk(job.mu);
        job.cv.wait(lk, [&] { return job.left.load() == 0; });
    }
};

// ------------------------------------------------------------------ file mappings
// A plain input file is mapped once and the mapping is kept for the next scan of the same file (same device, inode,
// size and mtime): DuckDB opens a reader at bind time and another one at init, a multi-GPU scan opens one per device,
// and a repeated query finds the page tables already populated (the first pass over a 2.8 GB tmpfs file spends ~20 % of
// its copy time in page faults, and unmapping it takes 25 ms).  At most four mappings are kept; evicted ones are unmapped
// off the critical path.
//
// Registered page cache (round 2).  After its first complete scan, a mapping that could be made writable-shared (the
// file opened O_RDWR; nothing is ever written) is registered with CUDA in the background (cudaHostRegister, portable):
// the file's PAGE-CACHE pages become pinned memory, and every later scan of the file copies them to the device by DMA
// straight from where they are -- no page cache -> pinned-block copy, which is what bounds a first scan (~46 GB/s on a
// 16-core host, and it does not scale with the GPUs of a box: profiles/round2_iobench2.txt).  Measured: registration
// 12 GB/s, H2D from the registered mapping 55.6 GB/s = the PCIe rate (profiles/round2_hostreg.txt).  Works for tmpfs and
// other memory-backed files; where the kernel refuses long-term pins (most disk file systems) registration fails and
// the copy path stays.  EXON_B200_REGISTER=0 turns it off, EXON_B200_REGISTER_MAX_GB caps the pinned page cache (32).
struct FileMap {
    const uint8_t* p = nullptr;
    int64_t size = 0;
    dev_t dev = 0;
    ino_t ino = 0;
    int64_t mtime_ns = 0;
    bool writable = false;            // mapped PROT_READ | PROT_WRITE, MAP_SHARED: the form cudaHostRegister accepts
    std::atomic<int> reg{0};          // 0 not registered, 1 registration running, 2 registered, 3 failed / not possible
    static std::atomic<int64_t>& registered_bytes() {
        static std::atomic<int64_t> b{0};
        return b;
    }
    // WHEN the mapping is registered matters: cudaHostRegister holds the driver's locks and the mapping's page locks for
    // as long as it pins pages (~0.1 s per GB), and a scan that runs meanwhile waits for it.  Measured on a 2.8 GB file
    // (profiles/round2_register_sequence.txt): registered right after the first scan's IO stage, that scan took 397
    // instead of 77 ms; registered in 64 MiB pieces regardless of activity, the next three scans took 433 / 257 / 126
    // instead of 65 ms; in pieces between scans, single scans stalled for up to 2 s.  So the registration is ONE call, made
    // once no reader of this process has been active for 200 ms -- the time between two statements of a session.  Under
    // back-to-back scans it does not get its turn and the copy path stays; a scan that starts while it runs waits for
    // it once.  (EXON_B200_REGISTER_PIECE_MB registers in pieces instead -- tests; a DMA never spans two pieces.)
    static constexpr int64_t REGISTER_MIN = 16 << 20;  // smaller files are not worth a registration (and are mapped read-only)
    static bool register_enabled() {
        static const bool on = !(getenv("EXON_B200_REGISTER") && atoi(getenv("EXON_B200_REGISTER")) == 0);
        return on;
    }
    // pinned page cache cannot be reclaimed: at most EXON_B200_REGISTER_MAX_GB (32) and a quarter of the host's memory
    static int64_t register_cap() {
        int64_t cap = (getenv("EXON_B200_REGISTER_MAX_GB") ? atoll(getenv("EXON_B200_REGISTER_MAX_GB")) : 32) << 30;
        if (FILE* f = fopen("/proc/meminfo", "r")) {
            long long kb = 0;
            if (fscanf(f, "MemTotal: %lld kB", &kb) == 1 && kb > 0) cap = std::min<int64_t>(cap, kb * 1024 / 4);
            fclose(f);
        }
        return cap;
    }
    static std::atomic<int>& io_active() {
        static std::atomic<int> n{0};
        return n;
    }
    int64_t piece = 1ll << 62;  // set when the registration starts, fixed from then on (default: the whole mapping is one piece)
    int64_t pieces_done = 0;   // (registration thread only, then the destructor)
    void unregister_pieces() {
        for (int64_t k = 0; k < pieces_done; k++) cudaHostUnregister(const_cast<uint8_t*>(p) + k * piece);
        pieces_done = 0;
    }
    ~FileMap() {
        if (p && reg.load() == 2) {
            unregister_pieces();
            registered_bytes().fetch_sub(size);
        }
        if (p) munmap(const_cast<uint8_t*>(p), (size_t)size);
    }
    // called when a scan has read the whole file once: register in the background (the scan that triggered it is done
    // with the host side; nobody waits for this)
    static void start_register(const std::shared_ptr<FileMap>& fm, int device) {
        static const int64_t cap = register_cap();
        if (!register_enabled() || !fm || !fm->writable || fm->size < REGISTER_MIN) return;
        int expect = 0;
        if (!fm->reg.compare_exchange_strong(expect, 1)) return;
        if (registered_bytes().fetch_add(fm->size) + fm->size > cap) {
            registered_bytes().fetch_sub(fm->size);
            fm->reg.store(3);
            return;
        }
        if (const char* e = getenv("EXON_B200_REGISTER_PIECE_MB")) fm->piece = (int64_t)std::max(1, atoi(e)) << 20;
        else fm->piece = std::max<int64_t>(fm->size, 1);
        std::thread([fm, device]() {
            cudaSetDevice(device);
            const int64_t pb = fm->piece;
            bool ok = true;
            // wait for 200 ms without an active reader (five minutes at most: then give up, a later scan may ask again)
            int idle = 0;
            for (int waited = 0; idle < 200 && waited < 300000; waited += 5) {
                std::this_thread::sleep_for(std::chrono::milliseconds(5));
                idle = io_active().load() > 0 ? 0 : idle + 5;
            }
            if (idle < 200) {
                registered_bytes().fetch_sub(fm->size);
                fm->reg.store(0);
                return;
            }
            for (int64_t off = 0; ok && off < fm->size; off += pb) {
                const int64_t len = std::min(pb, fm->size - off);
                if (cudaHostRegister(const_cast<uint8_t*>(fm->p) + off, (size_t)len, cudaHostRegisterPortable) == cudaSuccess) {
                    fm->pieces_done++;
                } else {
                    cudaGetLastError();
                    ok = false;
                }
            }
            if (ok) {
                fm->reg.store(2);
            } else {
                fm->unregister_pieces();
                registered_bytes().fetch_sub(fm->size);
                fm->reg.store(3);
            }
        }).detach();
    }
    // H2D copy out of a registered mapping: one DMA per piece the range touches
    static cudaError_t host_to_device(void* dst, const uint8_t* src, int64_t len, const FileMap* fm, cudaStream_t s) {
        if (!fm) return cudaMemcpyAsync(dst, src, (size_t)len, cudaMemcpyHostToDevice, s);
        const int64_t pb = fm->piece;
        int64_t off = src - fm->p;
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        while (len > 0) {
            const int64_t n = std::min(len, (off / pb + 1) * pb - off);
            const cudaError_t e = cudaMemcpyAsync(d, fm->p + off, (size_t)n, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return e;
            d += n;
            off += n;
            len -= n;
        }
        return cudaSuccess;
    }
};
struct MapCache {
    std::mutex mu;
    std::deque<std::shared_ptr<FileMap>> keep;
    static MapCache& get() {
        static MapCache* c = new MapCache();  // never destroyed (readers on detached threads may outlive static destructors)
        return *c;
    }
    // the cached mapping of a file that is already mapped, or nullptr (exb_file_cache_state)
    std::shared_ptr<FileMap> find(const struct stat& sb) {
        const int64_t mt = (int64_t)sb.st_mtim.tv_sec * 1000000000ll + sb.st_mtim.tv_nsec;
        std::lock_guard<std::mutex> lk(mu);
        for (auto& f : keep)
            if (f->dev == sb.st_dev && f->ino == sb.st_ino && f->size == (int64_t)sb.st_size && f->mtime_ns == mt) return f;
        return nullptr;
    }
    std::shared_ptr<FileMap> open(int fd, const char* path = nullptr) {
        struct stat sb;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size <= 0) return nullptr;
        const int64_t mt = (int64_t)sb.st_mtim.tv_sec * 1000000000ll + sb.st_mtim.tv_nsec;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (auto it = keep.begin(); it != keep.end(); ++it)
                if ((*it)->dev == sb.st_dev && (*it)->ino == sb.st_ino && (*it)->size == (int64_t)sb.st_size && (*it)->mtime_ns == mt) {
                    std::shared_ptr<FileMap> hit = *it;
                    keep.erase(it);
                    keep.push_front(hit);
                    return hit;
                }
        }
        // writable-shared if the file may be opened for writing (the form CUDA can register; nothing is written), else read-only
        void* m = MAP_FAILED;
        bool writable = false;
        if (path && FileMap::register_enabled() && (int64_t)sb.st_size >= FileMap::REGISTER_MIN) {
            const int wfd = ::open(path, O_RDWR);
            if (wfd >= 0) {
                struct stat wb;
                if (fstat(wfd, &wb) == 0 && wb.st_dev == sb.st_dev && wb.st_ino == sb.st_ino)
                    m = mmap(nullptr, (size_t)sb.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, wfd, 0);
                ::close(wfd);
                writable = m != MAP_FAILED;
            }
        }
        if (m == MAP_FAILED) m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) return nullptr;
        madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
        std::shared_ptr<FileMap> fm = std::make_shared<FileMap>();
        fm->writable = writable;
        fm->p = reinterpret_cast<const uint8_t*>(m);
        fm->size = sb.st_size;
        fm->dev = sb.st_dev;
        fm->ino = sb.st_ino;
        fm->mtime_ns = mt;
        std::shared_ptr<FileMap> evicted;
        {
            std::lock_guard<std::mutex> lk(mu);
            keep.push_front(fm);
            if (keep.size() > 4) {
                evicted = keep.back();
                keep.pop_back();
            }
        }
        if (evicted) std::thread([evicted]() mutable { evicted.reset(); }).detach();  // munmap of gigabytes: not on anybody's clock
        return fm;
    }
};

// ------------------------------------------------------------------ shard edges (include/exon_b200.h exb_reader_options)
// First record start at or after p.  Both ends of a byte-range shard go through the same function, so adjacent shards
// agree on the cut; a shard is then parsed as a complete input.
// FASTA: a line whose first byte is '>' always starts a record (noodles-fasta read_sequence stops there): exact.
static int64_t resync_fasta(const uint8_t* m, int64_t size, int64_t p) {
    if (p <= 0) return 0;
    while (p < size) {
        if (m[p] == '>' && m[p - 1] == '\n') return p;
        const void* q = memchr(m + p, '\n', (size_t)(size - p));
        if (!q) return size;
        p = (const uint8_t*)q - m + 1;
        if (p < size && m[p] == '>') return p;
    }
    return size;
}
// FASTQ: '@' and '+' are legal quality characters, so one line proves nothing.  Take the first line start >= p and the
// line starts after it; hypothesis h = "line h is a header" holds if every 4th line from h starts with '@' and the line
// two below it with '+', over a window of 16 records (a candidate needs its own four lines).  The smallest h that holds wins.  This is a heuristic for
// adversarial inputs; it cannot produce a silently wrong answer, because the shard that ENDS at the cut is parsed
// exactly from its own (inductively exact) start: if the cut is not a record boundary its line count is not a multiple
// of four or a header / plus line check fails, and the scan reports the malformed shard.
static int64_t resync_fastq(const uint8_t* m, int64_t size, int64_t p) {
    if (p <= 0) return 0;
    if (p >= size) return size;
    int64_t ls = p;
    if (m[p - 1] != '\n') {
        const void* q = memchr(m + p, '\n', (size_t)(size - p));
        if (!q) return size;
        ls = (const uint8_t*)q - m + 1;
    }
    constexpr int WINDOW = 4 * 16 + 4;
    int64_t starts[WINDOW];
    int L = 0;
    for (int64_t s = ls; s < size && L < WINDOW;) {
        starts[L++] = s;
        const void* q = memchr(m + s, '\n', (size_t)(size - s));
        if (!q) break;
        s = (const uint8_t*)q - m + 1;
    }
    // a record is four lines: with fewer than four lines left, no record starts at or after p
    if (L < 4) return size;
    for (int h = 0; h < 4 && h + 3 < L; h++) {
        bool ok = true;
        for (int j = h; j < L && ok; j += 4) {
            if (m[starts[j]] != '@') ok = false;
            else if (j + 2 < L && m[starts[j + 2]] != '+') ok = false;
        }
        if (ok) return starts[h];
    }
    return starts[0];  // nothing fits: the scan will report the malformed record
}

// ------------------------------------------------------------------ results
struct OutCol {  // one output column of the rows a chunk produced: views into the chunk's pinned result buffers
    bool present = false;
    int type = EXB_T_VARCHAR, kind = 0;
    const int64_t* off = nullptr;    // VARCHAR / list: rows + 1 entries starting at 0 (absent with EXB_RD_NO_OFFSETS)
    const uint8_t* data = nullptr;   // VARCHAR bytes
    const uint8_t* str = nullptr;    // rows x 16-byte string_t (EXB_RD_STRING_T)
    const uint8_t* valid = nullptr;  // one byte per row; nullptr = all valid
    const uint64_t* vbits = nullptr;
    const uint64_t* nulls_word = nullptr;  // number of NULL rows of the column in this chunk (valid after the D2H)
    const uint8_t* values = nullptr;     // numeric: rows x elem; list: child values (int32)
    int elem = 0;
    const uint64_t* entries = nullptr;   // list: rows x {offset, length}
    const int64_t* batch_base = nullptr; // list: first child value of batch k (batches of batch_size rows), closed with the total
};
constexpr int MAX_OUT = 4 + EXB_MAX_COMPUTED;
// The rows one input chunk produced.  Shared: every batch view handed out (exb_batch) holds a reference, so a
// host that borrows the strings (DuckDB string_t pointers) keeps the buffers alive past the reader's next step.
// The strings are handed out where the D2H copy put them -- no second host copy.
struct ChunkResult {
    OutCol cols[MAX_OUT];
    int64_t rows = 0;
    std::vector<HBuf*> bufs;
    std::shared_ptr<void> input;  // the pinned input block, when string_t entries of this chunk point into it
    std::shared_ptr<PinnedPool> pool;
    ~ChunkResult() {
        for (HBuf* b : bufs) pool->put(b);
    }
};

// One raw block of a file as the IO thread read it.  The part of the previous chunk that did not end on a record
// boundary never comes back to the host: it stays in HBM and the device thread puts the next block behind it.
// The raw bytes start BLOCK_SLACK bytes into the pinned buffer: the device thread copies the unconsumed tail of the
// previous chunk (the record that straddles two blocks) in front of them ON THE HOST too, so that the host holds a
// contiguous image of every chunk and string columns can point into it (string_t_borrow_kernel).
constexpr int64_t BLOCK_SLACK = 1 << 20;
struct Block {
    HBuf* h = nullptr;
    int64_t data_off = 0;      // raw bytes live at h->p + data_off
    const uint8_t* ext = nullptr;      // ... or at `ext`, inside a registered file mapping (h == nullptr): DMA straight from the page cache
    std::shared_ptr<FileMap> ext_map;  // keeps that mapping alive
    uint8_t* data() const { return ext ? const_cast<uint8_t*>(ext) : h->as<uint8_t>() + data_off; }
    // BGZF input inflated on the device (inflate.cu): data() holds zlen COMPRESSED bytes -- whole gzip members --, the
    // member table (n_z x exb_bgzf_block) sits at h->p + ztab_off, and raw_len is the size of their text
    int64_t zlen = 0, ztab_off = 0, n_z = 0, z_first = 0;
    int64_t raw_len = 0;
    int64_t raw_file_pos = 0;  // offset of the first raw byte in the (decompressed) file
    size_t file_idx = 0;
    bool eof = false;          // the file (or the shard of it) ends with this block
    std::string error;         // IO failure: the stream ends here
    bool end = false;          // no more files
};
struct HostBlock {  // an input block shared by the device thread and the chunk results whose strings point into it
    HBuf* h;                          // pinned block (nullptr when the bytes live in a registered file mapping)
    std::shared_ptr<PinnedPool> pool;
    std::shared_ptr<FileMap> map;     // ... which this keeps alive instead
    HostBlock(HBuf* b, std::shared_ptr<PinnedPool> p, std::shared_ptr<FileMap> m = nullptr) : h(b), pool(std::move(p)), map(std::move(m)) {}
    ~HostBlock() { pool->put(h); }
};
struct OutItem {
    std::shared_ptr<ChunkResult> res;
    int64_t counted = 0;
    bool end = false;
    std::string error;
};
template <typename T>
struct BoundedQueue {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<T> q;
    size_t cap;
    bool stop = false;
    explicit BoundedQueue(size_t c) : cap(c) {}
    bool push(T&& v) {  // false = stopped
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || q.size() < cap; });
        if (stop) return false;
        q.push_back(std::move(v));
        cv.notify_all();
        return true;
    }
    bool pop(T& out) {  // false = stopped
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || !q.empty(); });
        if (q.empty()) return false;
        out = std::move(q.front());
        q.pop_front();
        cv.notify_all();
        return true;
    }
    bool try_pop(T& out) {  // non-blocking: false = nothing queued right now
        std::lock_guard<std::mutex> lk(mu);
        if (q.empty()) return false;
        out = std::move(q.front());
        q.pop_front();
        cv.notify_all();
        return true;
    }
    void shutdown() {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
    }
};

struct NvtxRange {  // one range per pipeline stage (visible in nsys / ncu --nvtx timelines)
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------ the stream
// Three stages, each on its own thread, connected by bounded queues:
//   IO thread     file -> pinned blocks (plain files: mmap + memcpy by the IoPool workers; compressed: inflate)
//   device thread block -> H2D -> scan / filter / field split / gather / string_t kernels -> D2H into pinned result buffers
//   caller(s)     exb_reader_next / the Arrow stream: hand out <= batch_size-row views of a chunk's result
// so reading block k+1, the device work of block k, the D2H of block k-1 and the host's consumption overlap.
// ------------------------------------------------------------------ per-scan counters ("exon_gpu_stats()")
// Every reader leaves one record when it is closed; the last 64 are kept (exb_stats_snapshot, the extension's
// exon_gpu_stats() table function).
struct StatsRing {
    std::mutex mu;
    std::deque<exb_scan_stats> q;
    static StatsRing& get() {
        static StatsRing* r = new StatsRing();
        return *r;
    }
    void push(const exb_scan_stats& st) {
        std::lock_guard<std::mutex> lk(mu);
        q.push_front(st);
        if (q.size() > 64) q.pop_back();
    }
};

struct Reader {
    int format = 0;  // 1 FASTA, 2 FASTQ
    int ncols = 0;
    std::vector<std::string> col_names;
    int64_t batch_size = 2048;
    std::vector<std::string> files;
    std::vector<int> file_comp;
    int64_t chunk_bytes = 64ll << 20;
    // options
    int device = -1;
    int64_t range_lo = 0, range_hi = 0;  // byte-range shard (single plain file); range_hi <= 0: to the end
    uint32_t flags = 0;
    std::vector<exb_computed> computed;
    int n_out() const { return ncols + (int)computed.size(); }
    bool want_offsets() const { return !(flags & EXB_RD_NO_OFFSETS); }
    bool want_string_t() const { return (flags & EXB_RD_STRING_T) != 0; }
    // filter
    std::vector<Node> nodes;
    int root = -1;
    // device state (device thread only)
    cudaStream_t st = nullptr, sc = nullptr, sd = nullptr;  // compute stream, H2D prefetch stream, D2H stream
    cudaEvent_t ev_staged = nullptr, ev_stage_free = nullptr;
    DBuf d_inb[2], d_stage;  // the chunk being scanned / the one being assembled; the prefetched raw block
    // BGZF input: compressed members + their table on the device ([0] copied on the compute stream, [1] prefetched on
    // the copy stream), the inflate kernel's ticket + first-error word, and whether an inflate has run since the last check
    DBuf d_zin[2], d_zstate;
    bool z_pending = false;
    // compressed block -> device buffer `dz` (stream `s`): the members' bytes, then their table
    static int64_t z_table_at(const Block& b) { return (b.zlen + 16 + 63) & ~(int64_t)63; }
    bool z_copy_in(const Block& b, DBuf& dz, cudaStream_t s) {
        const int64_t tab_bytes = b.n_z * (int64_t)sizeof(exb_bgzf_block);
        if (!dz.need(z_table_at(b) + tab_bytes + 64)) return fail("out of device memory");
        return cu(FileMap::host_to_device(dz.p, b.data(), b.zlen, b.ext ? b.ext_map.get() : nullptr, s), "H2D compressed") &&
               cu(cudaMemcpyAsync(dz.as<uint8_t>() + z_table_at(b), b.h->as<uint8_t>() + b.ztab_off, (size_t)tab_bytes, cudaMemcpyHostToDevice, s), "H2D members");
    }
    // ... inflated into dst (stream `s`); the first-error word accumulates over the launches of a file
    bool z_inflate(const Block& b, DBuf& dz, uint8_t* dst, cudaStream_t s) {
        if (!d_zstate.p) {
            if (!d_zstate.need(16) || !cu(cudaMemsetAsync(d_zstate.p, 0xFF, 16, s), "memset")) return fail(derr.empty() ? "out of device memory" : derr);
        }
        z_pending = true;
        return cu(bgzf_inflate_launch(dz.as<uint8_t>(), reinterpret_cast<const exb_bgzf_block*>(dz.as<uint8_t>() + z_table_at(b)), b.n_z, dst,
                                      d_zstate.as<unsigned int>(), 1, b.z_first, s),
                  "bgzf_inflate");
    }
    // queue the read-back of the inflate kernel's error word with the chunk's other small results (slot 44) ...
    bool z_queue_status() { return !z_pending || queue_small(44, d_zstate.p, 2); }
    // ... and look at it once they are there
    bool z_check(const std::string& fname) {
        if (!z_pending) return true;
        z_pending = false;
        const unsigned int* w = reinterpret_cast<const unsigned int*>(small(44));
        int64_t bad = -1;
        if (exb_bgzf_status(w, &bad) == 0) return true;
        return fail(std::string(exb_last_error()) + " of " + fname);
    }
    uint8_t* d_cur = nullptr;  // = d_inb[cur].p while a chunk is processed
    DBuf d_ws, d_ws2, d_line, d_arr[4], d_lens, d_starts, d_valid, d_pass, d_selscratch, d_sel, d_lens2, d_starts2,
        d_valid2, d_off, d_cst, d_hdr_start, d_hdr_end, d_seq_off, d_gc_prefix, d_seq, d_err, d_info, d_qtmp, d_bad;
    // output buffers, double-buffered: the D2H of chunk k (stream sd) overlaps the kernels of chunk k+1 (stream st)
    struct OutSet {
        DBuf d_meta, d_data;
        cudaEvent_t ev_done = nullptr, ev_d2h = nullptr, ev_off = nullptr;
    } outs[2];
    int out_cur = 0;
    struct Pending {            // a result whose D2H is still in flight
        OutItem item;
        int set = -1;
        std::vector<std::pair<int, int>> bad;  // (out column, map mode) whose invalid-byte word must be checked after the D2H
        const uint64_t* bad_words = nullptr;
    };
    std::deque<Pending> pend;   // at most two: chunk k's D2H overlaps the kernels of chunk k+1
    std::shared_ptr<PinnedPool> pool = PinnedPool::shared();
    // a few words read back between launches (scan result, totals, counts): mapped pinned memory written by a kernel
    void* h_small = nullptr;   // host address
    void* d_small = nullptr;   // the same memory as the device sees it
    cudaEvent_t ev_small = nullptr;
    // rows ready to be handed out (caller threads, under call_mu)
    std::mutex call_mu;
    std::shared_ptr<ChunkResult> cur;
    int64_t rows = 0, next_row = 0;
    int64_t next_batch_index = 0;
    uint32_t column_mask = 0xF;          // bit c: column c is materialised (projection push-down)
    std::atomic<bool> count_only{false}; // COUNT(*): rows are counted, nothing is gathered or copied back
    int64_t counted = 0;
    int64_t rows_out = 0;                // rows that reached the caller (batches + counts): exon_gpu_stats
    std::atomic<int> any_bgzf{0}, any_direct{0};
    std::string error;  // caller's thread: the failure reported to the host
    std::string derr;   // device thread: failure of the chunk being processed
    // pipeline
    bool started = false, finished = false;
    std::thread io_thread, dev_thread;
    BoundedQueue<Block> inq{2};
    BoundedQueue<OutItem> outq{3};
    std::atomic<int64_t> block_bytes{0};
    std::atomic<bool> stopping{false};
    std::atomic<int64_t> bytes_done{0}, bytes_total{0};
    // device thread: the chunk being processed
    size_t cur_file = 0;
    int64_t cur_file_pos = 0;
    // EXON_B200_TRACE=1: seconds spent per stage, printed when the reader closes
    double t_io_read = 0, t_io_alloc = 0, t_io_push = 0, t_dev_pop = 0, t_dev_work = 0, t_dev_push = 0, t_call_pop = 0;
    double t_scan = 0, t_select = 0, t_mat = 0, t_flush = 0, t_init = 0, t_free = 0, t_map = 0, t_unmap = 0, t_first_block = 0, t_open = now();
    int64_t n_blocks = 0;
    int trace = 0;
    cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // EXON_B200_TRACE=2: GPU-side times per chunk
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    ~Reader() {
        stopping = true;
        inq.shutdown();
        outq.shutdown();
        if (io_thread.joinable()) io_thread.join();
        if (dev_thread.joinable()) dev_thread.join();
        for (Block& b : inq.q) pool->put(b.h);
        cur.reset();
        outq.q.clear();
        if (started) {
            exb_scan_stats st;
            memset(&st, 0, sizeof(st));
            snprintf(st.path, sizeof(st.path), "%s", files.empty() ? "" : files[0].c_str());
            st.n_files = (int32_t)files.size();
            st.format = format;
            st.device = device;
            st.compression = any_bgzf.load() ? 5 : (file_comp.empty() ? 0 : file_comp[0]);
            st.io_path = any_direct.load();
            st.file_bytes = bytes_total.load();
            st.bytes_done = bytes_done.load();
            st.rows = rows_out;
            st.blocks = n_blocks;
            st.seconds_total = now() - t_open;
            st.seconds_io = t_io_read;
            st.seconds_device = t_dev_work;
            st.seconds_scan = t_scan;
            st.seconds_select = t_select;
            st.seconds_materialise = t_mat;
            st.seconds_first_block = t_first_block;
            st.failed = error.empty() ? 0 : 1;
            StatsRing::get().push(st);
        }
        if (getenv("EXON_B200_TRACE"))
            fprintf(stderr, "exon_b200 reader: %lld blocks | io: alloc %.3f read %.3f push-wait %.3f | device: pop-wait %.3f work %.3f "
                            "(scan %.3f select %.3f materialise %.3f d2h-wait %.3f) push-wait %.3f | caller: pop-wait %.3f s | "
                            "device init %.3f free %.3f, map %.3f unmap %.3f, first block after %.3f, life %.3f s\n",
                    (long long)n_blocks, t_io_alloc, t_io_read, t_io_push, t_dev_pop, t_dev_work, t_scan, t_select, t_mat, t_flush, t_dev_push,
                    t_call_pop, t_init, t_free, t_map, t_unmap, t_first_block, now() - t_open);
    }
    bool fail(const std::string& m) {
        derr = m;
        return false;
    }
    bool cu(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        derr = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    }
    bool rc(int code) {
        if (code == 0) return true;
        derr = exb_last_error();
        return false;
    }
    // caller's thread: is there a device at all?  ("no CUDA device" is the message a CPU-only host must see, before any
    // pinned allocation is attempted)
    bool check_device() {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            cudaGetLastError();
            error = "no CUDA device: the exon_b200 scan engine has no CPU fallback";
            return false;
        }
        if (device < 0) {
            if (cudaGetDevice(&device) != cudaSuccess) device = 0;
        }
        if (device >= n) {
            error = "CUDA device " + std::to_string(device) + " does not exist (" + std::to_string(n) + " visible)";
            return false;
        }
        return true;
    }
    // device thread: streams and events live on THIS reader's device
    bool init_device() {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_staged, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_stage_free, cudaEventDisableTiming);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&outs[i].ev_done, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&outs[i].ev_d2h, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&outs[i].ev_off, cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaHostAlloc(&h_small, 4096, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d_small, h_small, 0);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_small, cudaEventDisableTiming);
        if (const char* t = getenv("EXON_B200_TRACE")) trace = atoi(t);
        if (trace >= 2)
            for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventCreate(&tev[i]);
        return cu(e, "CUDA stream setup");
    }
    void mark(int i) {
        if (trace >= 2) cudaEventRecord(tev[i], st);
    }
    // queue the read-back of `count` 8-byte words (d_src[i * stride]) into slot `slot` of the mapped buffer ...
    bool queue_small(int slot, const void* d_src, int count, int64_t stride = 1) {
        return cu(copy_words_launch(reinterpret_cast<uint8_t*>(d_small) + 8 * slot, d_src, count, stride, st), "copy_words");
    }
    // ... and wait for everything queued on the compute stream so far; the words are then readable at small(slot)
    bool wait_small() { return cu(cudaEventRecord(ev_small, st), "record") && cu(cudaEventSynchronize(ev_small), "sync"); }
    const int64_t* small(int slot) const { return reinterpret_cast<const int64_t*>(h_small) + slot; }
    void free_device() {  // device thread, at its end: buffers are freed on the device that owns them
        const double tf = now();
        for (DBuf* b : {&d_zin[0], &d_zin[1], &d_zstate, &d_inb[0], &d_inb[1], &d_stage, &d_ws, &d_ws2, &d_line, &d_arr[0], &d_arr[1], &d_arr[2], &d_arr[3], &d_lens, &d_starts,
                        &d_valid, &d_pass, &d_selscratch, &d_sel, &d_lens2, &d_starts2, &d_valid2, &d_off, &d_cst, &d_hdr_start, &d_hdr_end,
                        &d_seq_off, &d_gc_prefix, &d_seq, &d_err, &d_info, &d_qtmp, &d_bad, &outs[0].d_meta, &outs[0].d_data, &outs[1].d_meta,
                        &outs[1].d_data}) {
            b->release();
        }
        scratch.clear();
        for (int i = 0; i < 2; i++) {
            if (outs[i].ev_done) cudaEventDestroy(outs[i].ev_done);
            if (outs[i].ev_d2h) cudaEventDestroy(outs[i].ev_d2h);
            if (outs[i].ev_off) cudaEventDestroy(outs[i].ev_off);
        }
        if (ev_staged) cudaEventDestroy(ev_staged);
        if (ev_stage_free) cudaEventDestroy(ev_stage_free);
        if (ev_small) cudaEventDestroy(ev_small);
        if (h_small) cudaFreeHost(h_small);
        h_small = nullptr;
        if (st) cudaStreamDestroy(st);
        if (sc) cudaStreamDestroy(sc);
        if (sd) cudaStreamDestroy(sd);
        t_free += now() - tf;
    }

    // ------------------------------------------------------------------ IO thread
    static void pread_slices(int fd, uint8_t* dst, int64_t pos, int64_t want, int64_t* got_out, bool* err_out) {
        // fallback when the file cannot be mapped: ~5 GB/s per thread, so a block is read as eight slices
        constexpr int TMAX = 8;
        const int T = want >= (8ll << 20) ? TMAX : 1;
        int64_t got[TMAX] = {0};
        bool bad[TMAX] = {false};
        auto work = [&](int k) {
            const int64_t lo = want * k / T, hi = want * (k + 1) / T;
            int64_t done = 0;
            while (lo + done < hi) {
                const ssize_t r = pread(fd, dst + lo + done, (size_t)(hi - lo - done), (off_t)(pos + lo + done));
                if (r < 0) {
                    if (errno == EINTR) continue;
                    bad[k] = true;
                    break;
                }
                if (r == 0) break;
                done += r;
            }
            got[k] = done;
        };
        std::thread th[TMAX - 1];
        for (int k = 1; k < T; k++) th[k - 1] = std::thread(work, k);
        work(0);
        for (int k = 1; k < T; k++) th[k - 1].join();
        int64_t total = 0;
        *err_out = false;
        for (int k = 0; k < T; k++) {
            *err_out = *err_out || bad[k];
            total += got[k];
            if (got[k] < want * (k + 1) / T - want * k / T) break;  // short slice = end of file
        }
        *got_out = total;
    }
    void io_main() {
        Active active;
        cudaSetDevice(device);  // pinned allocations below belong to a context of this reader's device
        const bool sharded = range_lo > 0 || range_hi > 0;
        for (size_t fi = 0; fi < files.size() && !stopping; fi++) {
            const std::string& path = files[fi];
            const int comp = file_comp[fi];
            int fd = -1;
            gzFile gz = nullptr;
            FILE* zf = nullptr;  // zstd / bzip2 / xz: compressed file, decoder, its input window [zpos, zsize) of zbuf
            std::unique_ptr<Decoder> dec;
            std::vector<uint8_t> zbuf;
            size_t zpos = 0, zsize = 0;
            std::string err;
            std::shared_ptr<FileMap> fmap;  // plain files: the whole file mapped read-only (shared with later scans)
            const uint8_t* map = nullptr;
            int64_t map_size = 0;
            int64_t pos = 0, end_pos = -1;  // plain files: [pos, end_pos) is what this reader parses
            // BGZF (bgzip'ed) input: the compressed members go to the device as they are and are inflated there; the host
            // only walks their headers.  Any other gzip file takes the streaming zlib decoder below.
            bool bgzf = false;
            int64_t upos = 0, z_seen = 0;  // text bytes / members handed on so far
            std::vector<exb_bgzf_block> ztab;
            if (comp == 1 && !(getenv("EXON_B200_BGZF") && atoi(getenv("EXON_B200_BGZF")) == 0)) {
                fd = open(path.c_str(), O_RDONLY);
                if (fd >= 0) {
                    fmap = MapCache::get().open(fd, path.c_str());
                    if (fmap && exb_bgzf_probe_host(fmap->p, fmap->size)) {
                        bgzf = true;
                        map = fmap->p;
                        map_size = fmap->size;
                        end_pos = map_size;
                        ztab.resize(1 << 16);
                    } else {
                        fmap.reset();
                        close(fd);
                        fd = -1;
                    }
                }
            }
            if (bgzf) {
            } else if (comp == 1) {
                gz = gzopen(path.c_str(), "rb");
                if (!gz) err = "could not open " + path;
                else {
                    gzbuffer(gz, 1 << 20);
                    // the reference's GzipDecoder fails on input that is not gzip; gzread would pass it through
                    if (gzdirect(gz)) {
                        struct stat sb;
                        if (stat(path.c_str(), &sb) == 0 && sb.st_size > 0) err = "invalid gzip header in " + path;
                    }
                }
            } else if (comp == 0) {
                fd = open(path.c_str(), O_RDONLY);
                if (fd < 0) err = "could not open " + path;
                else {
                    const double tm = now();
                    fmap = MapCache::get().open(fd, path.c_str());
                    if (fmap) {
                        map = fmap->p;
                        map_size = fmap->size;
                    }
                    t_map += now() - tm;
                    if (sharded) {
                        if (!map) err = "byte-range shards need a mappable regular file: " + path;
                        else {
                            auto cut = [&](int64_t p) { return format == 1 ? resync_fasta(map, map_size, p) : resync_fastq(map, map_size, p); };
                            pos = cut(std::min(range_lo, map_size));
                            end_pos = range_hi > 0 ? cut(std::min(range_hi, map_size)) : map_size;
                            if (end_pos < pos) end_pos = pos;
                            shard_begin = pos;
                            shard_end = end_pos;
                        }
                    } else if (map) {
                        end_pos = map_size;
                    }
                }
            } else if (comp >= 2 && comp <= 4) {
                if (comp == 2) {
                    std::unique_ptr<ZstdDecoder> d(new ZstdDecoder());
                    if (d->init(err)) dec = std::move(d);
                } else if (comp == 3) {
                    std::unique_ptr<Bz2Decoder> d(new Bz2Decoder());
                    if (d->init(err)) dec = std::move(d);
                } else {
                    std::unique_ptr<XzDecoder> d(new XzDecoder());
                    if (d->init(err)) dec = std::move(d);
                }
                if (err.empty() && !(zf = fopen(path.c_str(), "rb"))) err = "could not open " + path;
                zbuf.resize(1 << 20);
            } else {
                err = "compression of " + path + " is not supported by this build";
            }
            bool eof = false;
            // DMA from the registered page cache (see FileMap) if an earlier scan of this file got it registered
            const bool direct = map && fmap && fmap->reg.load() == 2 && !(flags & EXB_RD_COPY_IO);
            io_direct.store(direct ? 1 : 0);
            if (direct) any_direct.store(1);
            if (bgzf) any_bgzf.store(1);
            while (err.empty() && !eof && !stopping) {
                // block edges sit on multiples of 16 in file coordinates (the first block of a shard may be a little shorter):
                // the chained COUNT scan of dev_main_fused needs that of every range but the first
                const int64_t want = block_bytes.load() - (pos & 15);
                Block b;
                double t0 = now();
                if (bgzf) {
                    int64_t nb = 0, next = pos, outb = 0;
                    // (text per block: 4 x the plain block size -- one warp inflates one member, and a launch wants thousands of them)
                    static const int64_t zmul = getenv("EXON_B200_BGZF_BLOCK_MUL") ? std::max(1, atoi(getenv("EXON_B200_BGZF_BLOCK_MUL"))) : 4;
                    if (exb_bgzf_index_host(map, map_size, pos, block_bytes.load() * zmul, ztab.data(), (int64_t)ztab.size(), &nb, &next, &outb) != 0) {
                        err = std::string(exb_last_error()) + " in " + path;
                        break;
                    }
                    const int64_t zlen = next - pos, tab_bytes = nb * (int64_t)sizeof(exb_bgzf_block);
                    const int64_t zoff = direct ? 0 : 64, toff = direct ? 0 : ((zoff + zlen + 63) & ~(int64_t)63);
                    b.h = pool->get(toff + tab_bytes + 64);
                    if (!b.h) {
                        err = "out of pinned host memory";
                        break;
                    }
                    if (direct) {  // the compressed bytes are DMAed straight from the registered page cache
                        b.ext = map + pos;
                        b.ext_map = fmap;
                    } else {
                        b.data_off = zoff;
                        IoPool::get().copy(b.h->as<uint8_t>() + zoff, map + pos, zlen);
                    }
                    memcpy(b.h->as<uint8_t>() + toff, ztab.data(), (size_t)tab_bytes);
                    b.zlen = zlen;
                    b.ztab_off = toff;
                    b.n_z = nb;
                    b.z_first = z_seen;
                    z_seen += nb;
                    b.file_idx = fi;
                    b.raw_file_pos = upos;
                    b.raw_len = outb;
                    upos += outb;
                    pos = next;
                    eof = pos >= map_size;
                    b.eof = eof;
                    t_io_read += now() - t0;
                    if (n_blocks == 0) t_first_block = now() - t_open;
                    n_blocks++;
                    HBuf* hb = b.h;
                    if (!inq.push(std::move(b))) {
                        pool->put(hb);
                        break;
                    }
                    continue;
                }
                if (direct) {  // the mapping is pinned: the block IS the file's page cache, nothing to copy
                    const int64_t got_d = std::min<int64_t>(want, end_pos - pos);
                    b.ext = map + pos;
                    b.ext_map = fmap;
                    b.file_idx = fi;
                    b.raw_file_pos = pos;
                    b.raw_len = got_d;
                    eof = pos + got_d >= end_pos;
                    b.eof = eof;
                    pos += got_d;
                    if (n_blocks == 0) t_first_block = now() - t_open;
                    n_blocks++;
                    if (!inq.push(std::move(b))) break;
                    continue;
                }
                b.h = pool->get(BLOCK_SLACK + want + 64);
                b.data_off = BLOCK_SLACK;
                t_io_alloc += now() - t0;
                if (!b.h) {
                    err = "out of pinned host memory";
                    break;
                }
                t0 = now();
                NvtxRange nv("exb:io_block");
                b.file_idx = fi;
                b.raw_file_pos = pos;
                uint8_t* dst = b.data();
                int64_t got = 0;
                if (gz) {
                    while (got < want) {
                        const int g = gzread(gz, dst + got, (unsigned)std::min<int64_t>(want - got, 1 << 30));
                        if (g < 0) {
                            err = "gzip read error in " + path;
                            break;
                        }
                        if (g == 0) {
                            // a truncated stream makes gzread return the partial data and then 0 with only gzerror() set
                            int zerr = Z_OK;
                            gzerror(gz, &zerr);
                            if (zerr == Z_BUF_ERROR || zerr == Z_DATA_ERROR) err = "truncated or corrupt gzip stream in " + path;
                            break;
                        }
                        got += g;
                    }
                } else if (dec) {
                    while (got < want) {
                        if (zpos == zsize) {
                            const size_t n = fread(zbuf.data(), 1, zbuf.size(), zf);
                            if (n == 0) {
                                if (ferror(zf)) err = "read error in " + path;
                                else if (dec->in_stream()) err = std::string("truncated ") + dec->name() + " stream in " + path;
                                break;  // end of the compressed file
                            }
                            zpos = 0;
                            zsize = n;
                        }
                        size_t opos = (size_t)got;
                        const size_t ipos0 = zpos;
                        if (!dec->step(zbuf.data(), zsize, zpos, dst, (size_t)want, opos, err)) {
                            err += " in " + path;
                            break;
                        }
                        if (zpos == ipos0 && opos == (size_t)got && zpos < zsize) {  // no progress with input and room available
                            err = std::string(dec->name()) + ": decoder stalled in " + path;
                            break;
                        }
                        got = (int64_t)opos;
                    }
                } else if (map) {
                    got = std::min<int64_t>(want, end_pos - pos);
                    IoPool::get().copy(dst, map + pos, got);
                } else if (fd >= 0) {
                    bool bad = false;
                    pread_slices(fd, dst, pos, want, &got, &bad);
                    if (bad) err = "read error in " + path;
                }
                if (!err.empty()) {
                    pool->put(b.h);
                    break;
                }
                t_io_read += now() - t0;
                if (n_blocks == 0) t_first_block = now() - t_open;
                n_blocks++;
                eof = map ? pos + got >= end_pos : got < want;
                b.raw_len = got;
                b.eof = eof;
                pos += got;
                HBuf* hb = b.h;
                t0 = now();
                const bool pushed = inq.push(std::move(b));
                t_io_push += now() - t0;
                if (!pushed) {  // reader closed
                    pool->put(hb);
                    break;
                }
            }
            // the whole file went by once through the copy path: have its page cache registered for the next scan
            if (map && fmap && err.empty() && !stopping && !direct && !(flags & EXB_RD_COPY_IO)) FileMap::start_register(fmap, device);
            fmap.reset();
            if (fd >= 0) close(fd);
            if (gz) gzclose(gz);
            dec.reset();
            if (zf) fclose(zf);
            if (!err.empty()) {
                Block b;
                b.error = err;
                b.end = true;
                inq.push(std::move(b));
                return;
            }
        }
        Block b;
        b.end = true;
        inq.push(std::move(b));
    }
    int64_t shard_begin = 0, shard_end = 0;  // where the byte-range shard really starts / ends (after resync)
    // host image of the chunk being processed: host_base[i] is the byte d_cur[i] (nullptr: no contiguous image, see
    // dev_main); host_block keeps the pinned block that holds it alive
    std::atomic<int> io_direct{0};  // the current file is read by DMA from its registered page cache (exb_reader_io_path)
    uint8_t* host_base = nullptr;
    std::shared_ptr<HostBlock> host_block;
    bool no_borrow = getenv("EXON_B200_NO_BORROW") != nullptr;  // debugging / A-B: always gather and copy the strings back

    // ------------------------------------------------------------------ device thread
    // evaluate node `k` into d_out (uint8 per row) using per-column starts/lens (cols x n) in column buffers
    struct EvalCtx {
        const uint8_t* const* col_buf;  // per column: base buffer the starts index into
        const int64_t* starts;          // [ncols][n]
        const uint32_t* lens;           // [ncols][n]
        const uint8_t* desc_valid;
        int64_t n;
    };
    bool eval(int k, const EvalCtx& c, uint8_t* d_out, int depth) {
        const Node& nd = nodes[k];
        if (nd.kind == N_AND || nd.kind == N_OR) {
            if (!eval(nd.lhs, c, d_out, depth)) return false;
            // right operand into a scratch of its own (one per nesting level)
            if ((int)scratch.size() <= depth) scratch.resize(depth + 1);
            if (!scratch[depth]) scratch[depth].reset(new DBuf());
            if (!scratch[depth]->need(c.n)) return fail("out of device memory");
            uint8_t* tmp = scratch[depth]->as<uint8_t>();
            if (!eval(nd.rhs, c, tmp, depth + 1)) return false;
            return cu(pass_combine_launch(d_out, tmp, c.n, nd.kind == N_AND ? 0 : 1, st), "pass_combine");
        }
        if (nd.kind == N_STR) {
            if (!d_cst.need((int64_t)nd.cst.size() + 16)) return fail("out of device memory");
            // constants are tiny; the copy is stream ordered with the kernel that reads it
            cst_keep.push_back(nd.cst);
            if (!nd.cst.empty() &&
                !cu(cudaMemcpyAsync(d_cst.p, cst_keep.back().data(), nd.cst.size(), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(const)"))
                return false;
            const bool is_desc = col_names[nd.col] == "description";
            bool ok = cu(str_pred_launch(c.col_buf[nd.col], c.starts + (int64_t)nd.col * c.n, c.lens + (int64_t)nd.col * c.n,
                                         is_desc ? c.desc_valid : nullptr, c.n, nd.op, d_cst.as<uint8_t>(), (int)nd.cst.size(), d_out, st),
                         "str_pred");
            // d_cst is reused by the next leaf: keep launches and copies ordered
            return ok && cu(cudaStreamSynchronize(st), "sync");
        }
        // numeric
        if (format == 2) {
            exb_predicate p{nd.field, nd.op, nd.value};
            if (!d_selscratch.need(64)) return fail("out of device memory");
            return rc(exb_fastq_filter(d_arr[0].as<uint32_t>(), d_arr[1].as<uint32_t>(), d_arr[2].as<uint32_t>(), d_arr[3].as<int32_t>(), c.n,
                                       &p, 1, d_out, d_selscratch.as<int64_t>(), nullptr, st));
        }
        if (nd.field != EXB_P_GC_CONTENT && nd.field != EXB_P_SEQ_LEN) return fail("predicate not applicable to FASTA");
        return cu(fasta_num_pred_launch(d_seq_off.as<int64_t>(), d_gc_prefix.as<int64_t>(), c.n, nd.field, nd.op, nd.value, d_out, st),
                  "fasta_num_pred");
    }
    std::vector<std::unique_ptr<DBuf>> scratch;
    std::vector<std::string> cst_keep;

    bool needs_numeric() const {
        for (const Node& n : nodes)
            if (n.kind == N_NUM) return true;
        return false;
    }
    bool has_computed(int kind) const {
        for (const exb_computed& c : computed)
            if (c.kind == kind) return true;
        return false;
    }

    // hand finished results to the caller side once their D2H has drained, oldest first, until at most `keep` are in flight
    bool flush_pending(size_t keep) {
        while (pend.size() > keep) {
            Pending pd = std::move(pend.front());
            pend.pop_front();
            const double t0 = now();
            OutSet& o = outs[pd.set];
            const bool ok = cu(cudaEventSynchronize(o.ev_d2h), "D2H") && cu(cudaEventSynchronize(o.ev_off), "D2H offsets");
            t_flush += now() - t0;
            if (!ok) return false;
            for (auto& pb : pd.bad) {  // reverse_complement & co met a byte outside their table (module.cpp:58-62)
                const uint64_t w = pd.bad_words[pb.first];
                if (w != ~0ull) return fail(std::string("Invalid character in sequence: ") + std::string(1, (char)(w & 0xFF)));
            }
            const double t1 = now();
            const bool pushed = outq.push(std::move(pd.item));
            t_dev_push += now() - t1;
            if (!pushed) {
                stopping = true;
                pend.clear();
                return true;
            }
        }
        return true;
    }

    // Gather the selected rows of the wanted columns, build what the host's vectors hold (string_t entries, validity
    // bitmaps, computed columns) and bring it to the host: ONE launch for the Arrow offsets of all file columns, one
    // gather per wanted column into one device buffer, then metadata and bytes cross PCIe in two copies on the D2H
    // stream into pinned buffers that the batches hand out as they are.  One stream sync (the column totals).
    //   col_buf / d_st / d_ln / d_val: source buffer, start, length, description validity of the n selected rows
    //   sel: the selected records (nullptr = all), for the per-record arrays the computed columns read
    bool materialise(const uint8_t* const* col_buf, const int64_t* d_st, const uint32_t* d_ln, const uint8_t* d_val, const int64_t* sel, int64_t n,
                     OutItem* item) {
        if (count_only.load()) {
            item->counted = n;
            return true;
        }
        if (n == 0) return true;  // nothing to hand out for this chunk
        NvtxRange nv("exb:materialise");
        const int seq_col = 2, qual_col = 3;
        const int nout = n_out();
        // which file column feeds output column j (-1 = none), and how
        int src[MAX_OUT], type[MAX_OUT], kind[MAX_OUT], mode[MAX_OUT];
        bool present[MAX_OUT];
        for (int j = 0; j < nout; j++) {
            src[j] = -1;
            mode[j] = -1;
            kind[j] = 0;
            type[j] = EXB_T_VARCHAR;
            if (j < ncols) {
                present[j] = ((column_mask >> j) & 1u) != 0;
                src[j] = j;
                continue;
            }
            const exb_computed& c = computed[j - ncols];
            present[j] = true;
            kind[j] = c.kind;
            switch (c.kind) {
            case EXB_C_GC_CONTENT: type[j] = EXB_T_FLOAT; break;
            case EXB_C_SEQ_MAP: type[j] = EXB_T_VARCHAR; src[j] = seq_col; mode[j] = c.arg; break;
            case EXB_C_QUALITY_LIST: type[j] = EXB_T_INT32_LIST; src[j] = qual_col; break;
            case EXB_C_MEAN_QUALITY: type[j] = EXB_T_DOUBLE; break;
            case EXB_C_SEQ_LENGTH: type[j] = EXB_T_INT64; src[j] = seq_col; break;
            case EXB_C_QUAL_LENGTH: type[j] = EXB_T_INT64; src[j] = qual_col; break;
            default: return fail("internal: unknown computed column");
            }
        }
        OutSet& o = outs[out_cur];
        const int set = out_cur;
        out_cur ^= 1;
        std::shared_ptr<ChunkResult> res = std::make_shared<ChunkResult>();
        res->pool = pool;
        res->rows = n;
        // ---- borrowed columns: a VARCHAR column whose rows are byte ranges of the input chunk (every FASTQ column, FASTA id /
        // description) is not gathered at all when the host reads string_t entries only and holds a contiguous image of the
        // chunk: the entries point into the pinned INPUT block (string_t_borrow_kernel), which the result keeps alive.
        bool borrow[MAX_OUT];
        bool need_offsets = false;
        for (int j = 0; j < nout; j++) {
            borrow[j] = present[j] && type[j] == EXB_T_VARCHAR && mode[j] < 0 && want_string_t() && !want_offsets() && host_base != nullptr &&
                        col_buf[src[j]] == d_cur && !no_borrow;
        }
        {
            // ONE long column on its own (sequence or quality_scores, ~40 % of a record) is still gathered: a consumer that
            // reads the bytes then walks a dense, freshly written buffer instead of 150 bytes of every 360 of the input
            // image (DuckDB's SUM(length(sequence)): 35 GB/s gathered, 24 GB/s borrowed; all four columns: 11.5 -> 24 GB/s)
            int n_borrow = 0, last = -1;
            for (int j = 0; j < nout; j++)
                if (borrow[j]) {
                    n_borrow++;
                    last = j;
                }
            if (n_borrow == 1 && src[last] >= 2 && !getenv("EXON_B200_BORROW_ALWAYS")) borrow[last] = false;
        }
        for (int j = 0; j < nout; j++) {
            if (present[j] && ((type[j] == EXB_T_VARCHAR && !borrow[j]) || type[j] == EXB_T_INT32_LIST)) need_offsets = true;
        }
        int64_t* d_offs = nullptr;
        int64_t totals[4] = {0, 0, 0, 0};
        if (need_offsets) {  // (one multi-column scan + a stream sync for the totals: skipped when every string column is borrowed)
            const int64_t ws_bytes = exb_scan_workspace_bytes(4 * n * ncols + 16);
            if (!d_ws2.need(ws_bytes) || !d_off.need((int64_t)ncols * (n + 1) * 8)) return fail("out of memory");
            d_offs = d_off.as<int64_t>();
            if (!rc(exb_exclusive_scan_u32_multi(d_ln, n, ncols, n, d_offs, n + 1, d_ws2.p, d_ws2.cap, st))) return false;
            if (!queue_small(32, d_offs + n, ncols, n + 1) || !wait_small()) return false;
            for (int c = 0; c < ncols; c++) totals[c] = small(32)[c];
        }
        // ---- layout of the two result buffers
        auto up = [](int64_t x, int64_t a) { return (x + a - 1) & ~(a - 1); };
        const int64_t n_batches = (n + batch_size - 1) / batch_size;
        const int64_t vwords = (n + 63) / 64;
        int64_t data_base[MAX_OUT] = {0}, all = 0;
        int64_t m = 0;  // metadata cursor
        const int64_t m_nulls = m;  m += up((int64_t)MAX_OUT * 8, 64);   // null count per column
        const int64_t m_bad = m;    m += up((int64_t)MAX_OUT * 8, 64);   // invalid-byte word per mapped column
        int64_t m_str[MAX_OUT], m_valid[MAX_OUT], m_vbits[MAX_OUT], m_values[MAX_OUT], m_entries[MAX_OUT], m_bases[MAX_OUT];
        for (int j = 0; j < nout; j++) {
            m_str[j] = m_valid[j] = m_vbits[j] = m_values[j] = m_entries[j] = m_bases[j] = -1;
            if (!present[j]) continue;
            if (type[j] == EXB_T_VARCHAR) {
                data_base[j] = all;
                if (!borrow[j]) all += up(totals[src[j]], 16);
                if (want_string_t()) { m_str[j] = m; m += up(n * 16, 64); }
                if (j < ncols && col_names[j] == "description") {
                    m_valid[j] = m; m += up(n, 64);
                    m_vbits[j] = m; m += up(vwords * 8, 64);
                }
            } else if (type[j] == EXB_T_INT32_LIST) {
                data_base[j] = all;
                all += up(4 * totals[src[j]], 16);
                m_entries[j] = m; m += up(n * 16, 64);
                m_bases[j] = m;   m += up((n_batches + 1) * 8, 64);
            } else {
                const int elem = type[j] == EXB_T_FLOAT ? 4 : 8;
                m_values[j] = m; m += up(n * elem, 64);
                if (kind[j] == EXB_C_MEAN_QUALITY) {
                    m_valid[j] = m; m += up(n, 64);
                    m_vbits[j] = m; m += up(vwords * 8, 64);
                }
            }
        }
        const int64_t meta_copy = m;  // [0, meta_copy) of the device metadata buffer is mirrored on the host in one copy
        int64_t m_off[MAX_OUT];       // Arrow offsets: copied column by column from the scan's output, behind the mirror
        for (int j = 0; j < nout; j++) {
            m_off[j] = -1;
            if (present[j] && want_offsets() && (type[j] == EXB_T_VARCHAR || type[j] == EXB_T_INT32_LIST)) { m_off[j] = m; m += up((n + 1) * 8, 64); }
        }
        HBuf* h_meta = pool->get(m + 64);
        HBuf* h_bytes = pool->get(all + 64);
        if (h_meta) res->bufs.push_back(h_meta);
        if (h_bytes) res->bufs.push_back(h_bytes);
        if (!h_meta || !h_bytes || !o.d_meta.need(meta_copy + 64) || !o.d_data.need(all + 64)) return fail("out of memory");
        // the set's previous D2H (two chunks ago) was waited for by flush_pending(1) after the last chunk; order the device side too
        if (!cu(cudaStreamWaitEvent(st, o.ev_d2h, 0), "wait")) return false;
        uint8_t* dm = o.d_meta.as<uint8_t>();
        uint8_t* dd = o.d_data.as<uint8_t>();
        uint8_t* hm = h_meta->as<uint8_t>();
        uint8_t* hb = h_bytes->as<uint8_t>();
        if (!cu(cudaMemsetAsync(dm + m_nulls, 0, (size_t)(m_bad - m_nulls), st), "memset")) return false;
        if (!cu(cudaMemsetAsync(dm + m_bad, 0xFF, (size_t)up((int64_t)MAX_OUT * 8, 64), st), "memset")) return false;
        std::vector<std::pair<int, int>> bad_cols;
        for (int j = 0; j < nout; j++) {
            if (!present[j]) continue;
            OutCol& oc = res->cols[j];
            oc.present = true;
            oc.nulls_word = reinterpret_cast<const uint64_t*>(hm + m_nulls) + j;
            oc.type = type[j];
            oc.kind = kind[j];
            const int c = src[j];
            if (type[j] == EXB_T_VARCHAR && borrow[j]) {
                if (!cu(string_t_borrow_launch(d_st + (int64_t)c * n, d_ln + (int64_t)c * n, d_cur, (uint64_t)(uintptr_t)host_base, n, dm + m_str[j], st),
                        "string_t_borrow"))
                    return false;
                oc.str = hm + m_str[j];
                res->input = host_block;
                if (m_valid[j] >= 0) {
                    if (!cu(cudaMemcpyAsync(dm + m_valid[j], d_val, (size_t)n, cudaMemcpyDeviceToDevice, st), "D2D validity")) return false;
                    if (!cu(valid_bits_launch(d_val, n, reinterpret_cast<uint64_t*>(dm + m_vbits[j]), reinterpret_cast<uint64_t*>(dm + m_nulls) + j, st),
                            "valid_bits"))
                        return false;
                    oc.valid = hm + m_valid[j];
                    oc.vbits = reinterpret_cast<const uint64_t*>(hm + m_vbits[j]);
                }
            } else if (type[j] == EXB_T_VARCHAR) {
                const int64_t* offs = d_offs + (int64_t)c * (n + 1);
                if (totals[c] > 0) {
                    if (mode[j] < 0) {
                        if (!rc(exb_gather_ranges(col_buf[c], d_st + (int64_t)c * n, d_ln + (int64_t)c * n, offs, n, dd + data_base[j], totals[c], st)))
                            return false;
                    } else {
                        if (!cu(gather_ranges_map_launch(col_buf[c], d_st + (int64_t)c * n, offs, n, totals[c], dd + data_base[j], mode[j],
                                                         reinterpret_cast<unsigned long long*>(dm + m_bad) + j, st),
                                "gather_map"))
                            return false;
                        bad_cols.emplace_back(j, mode[j]);
                    }
                }
                oc.data = hb + data_base[j];
                if (m_str[j] >= 0) {
                    if (!cu(string_t_launch(offs, dd + data_base[j], (uint64_t)(uintptr_t)(hb + data_base[j]), n, dm + m_str[j], st), "string_t")) return false;
                    oc.str = hm + m_str[j];
                }
                if (m_valid[j] >= 0) {
                    if (!cu(cudaMemcpyAsync(dm + m_valid[j], d_val, (size_t)n, cudaMemcpyDeviceToDevice, st), "D2D validity")) return false;
                    if (!cu(valid_bits_launch(d_val, n, reinterpret_cast<uint64_t*>(dm + m_vbits[j]), reinterpret_cast<uint64_t*>(dm + m_nulls) + j, st),
                            "valid_bits"))
                        return false;
                    oc.valid = hm + m_valid[j];
                    oc.vbits = reinterpret_cast<const uint64_t*>(hm + m_vbits[j]);
                }
            } else if (type[j] == EXB_T_INT32_LIST) {
                const int64_t* offs = d_offs + (int64_t)c * (n + 1);
                if (totals[c] > 0) {
                    if (!d_qtmp.need(totals[c] + 64)) return fail("out of device memory");
                    if (!rc(exb_gather_ranges(col_buf[c], d_st + (int64_t)c * n, d_ln + (int64_t)c * n, offs, n, d_qtmp.as<uint8_t>(), totals[c], st)))
                        return false;
                    if (!rc(exb_quality_decode(d_qtmp.as<uint8_t>(), totals[c], reinterpret_cast<int32_t*>(dd + data_base[j]), st))) return false;
                }
                if (!cu(list_entries_launch(offs, n, batch_size, dm + m_entries[j], reinterpret_cast<int64_t*>(dm + m_bases[j]), st), "list_entries"))
                    return false;
                oc.values = hb + data_base[j];
                oc.elem = 4;
                oc.entries = reinterpret_cast<const uint64_t*>(hm + m_entries[j]);
                oc.batch_base = reinterpret_cast<const int64_t*>(hm + m_bases[j]);
            } else {
                cudaError_t e = cudaSuccess;
                if (kind[j] == EXB_C_GC_CONTENT) {
                    if (format == 2)
                        e = gc_sel_launch(d_arr[0].as<uint32_t>(), d_arr[1].as<uint32_t>(), nullptr, nullptr, sel, n, reinterpret_cast<float*>(dm + m_values[j]), st);
                    else
                        e = gc_sel_launch(nullptr, nullptr, d_seq_off.as<int64_t>(), d_gc_prefix.as<int64_t>(), sel, n, reinterpret_cast<float*>(dm + m_values[j]), st);
                    oc.elem = 4;
                } else if (kind[j] == EXB_C_MEAN_QUALITY) {
                    e = mean_quality_launch(d_arr[2].as<uint32_t>(), d_arr[3].as<int32_t>(), sel, n, reinterpret_cast<double*>(dm + m_values[j]),
                                            dm + m_valid[j], st);
                    if (e == cudaSuccess)
                        e = valid_bits_launch(dm + m_valid[j], n, reinterpret_cast<uint64_t*>(dm + m_vbits[j]), reinterpret_cast<uint64_t*>(dm + m_nulls) + j, st);
                    oc.elem = 8;
                    oc.valid = hm + m_valid[j];
                    oc.vbits = reinterpret_cast<const uint64_t*>(hm + m_vbits[j]);
                } else {
                    e = lens_i64_launch(d_ln + (int64_t)c * n, n, reinterpret_cast<int64_t*>(dm + m_values[j]), st);
                    oc.elem = 8;
                }
                if (!cu(e, "computed column")) return false;
                oc.values = hm + m_values[j];
            }
            if (m_off[j] >= 0) {
                if (!cu(cudaMemcpyAsync(hm + m_off[j], d_offs + (int64_t)c * (n + 1), (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H offsets")) return false;
                oc.off = reinterpret_cast<const int64_t*>(hm + m_off[j]);
            }
        }
        if (!cu(cudaEventRecord(o.ev_off, st), "record") || !cu(cudaEventRecord(o.ev_done, st), "record")) return false;
        // metadata + bytes on the D2H stream: the next chunk's kernels do not wait for PCIe
        if (!cu(cudaStreamWaitEvent(sd, o.ev_done, 0), "wait")) return false;
        if (!cu(cudaMemcpyAsync(hm, dm, (size_t)meta_copy, cudaMemcpyDeviceToHost, sd), "D2H metadata")) return false;
        if (all > 0 && !cu(cudaMemcpyAsync(hb, dd, (size_t)all, cudaMemcpyDeviceToHost, sd), "D2H data")) return false;
        if (!cu(cudaEventRecord(o.ev_d2h, sd), "record")) return false;
        Pending pd;
        pd.item.res = res;
        pd.set = set;
        pd.bad = bad_cols;
        pd.bad_words = reinterpret_cast<const uint64_t*>(hm + m_bad);
        pend.push_back(std::move(pd));
        // deliver chunk k-1 now: its D2H had all of chunk k's kernels to finish, and chunk k's is already queued behind it
        return flush_pending(1);
    }

    // apply the filter (if any) to the n rows described by starts/lens/valid; leaves the final arrays in *o_*
    bool select(const uint8_t* const* col_buf, int64_t n, const int64_t*& o_st, const uint32_t*& o_ln, const uint8_t*& o_val, const int64_t*& o_sel,
                int64_t& o_n) {
        o_st = d_starts.as<int64_t>();
        o_ln = d_lens.as<uint32_t>();
        o_val = d_valid.as<uint8_t>();
        o_sel = nullptr;
        o_n = n;
        if (root < 0 || n == 0) return true;
        NvtxRange nv("exb:select");
        if (!d_pass.need(n)) return fail("out of device memory");
        EvalCtx c{col_buf, o_st, o_ln, o_val, n};
        cst_keep.clear();
        if (!eval(root, c, d_pass.as<uint8_t>(), 0)) return false;
        const int64_t ws_bytes = exb_scan_workspace_bytes(4 * n + 16);
        if (!d_ws2.need(ws_bytes) || !d_off.need((n + 1) * 8) || !d_sel.need(n * 8)) return fail("out of device memory");
        if (!rc(exb_select_rows(d_pass.as<uint8_t>(), n, d_off.as<int64_t>(), d_sel.as<int64_t>(), d_ws2.p, d_ws2.cap, st))) return false;
        if (!queue_small(40, d_off.as<int64_t>() + n, 1) || !wait_small()) return false;
        const int64_t cnt = small(40)[0];
        o_n = cnt;
        o_sel = d_sel.as<int64_t>();
        if (count_only.load()) return true;  // COUNT(*): the number of passing rows is all that is needed
        if (!d_lens2.need(std::max<int64_t>(cnt, 1) * 4 * ncols) || !d_starts2.need(std::max<int64_t>(cnt, 1) * 8 * ncols) ||
            !d_valid2.need(std::max<int64_t>(cnt, 1)))
            return fail("out of device memory");
        for (int k = 0; k < ncols; k++) {
            if (!cu(take_u32_launch(o_ln + (int64_t)k * n, d_sel.as<int64_t>(), cnt, d_lens2.as<uint32_t>() + (int64_t)k * cnt, st), "take")) return false;
            if (!cu(take_i64_launch(o_st + (int64_t)k * n, d_sel.as<int64_t>(), cnt, d_starts2.as<int64_t>() + (int64_t)k * cnt, st), "take")) return false;
        }
        if (!cu(take_u8_launch(o_val, d_sel.as<int64_t>(), cnt, d_valid2.as<uint8_t>(), st), "take")) return false;
        o_st = d_starts2.as<int64_t>();
        o_ln = d_lens2.as<uint32_t>();
        o_val = d_valid2.as<uint8_t>();
        return true;
    }

    // process the n bytes at d_cur (already enqueued on `st`); sets `consumed`.  Returns false on error (derr).
    bool process_chunk(int64_t n, bool is_final, int64_t& consumed, bool& grew, OutItem* item) {
        grew = false;
        const std::string& fname = files[cur_file];
        const int64_t ws_bytes = exb_scan_workspace_bytes(n + 16);
        if (!d_ws.need(ws_bytes) || !d_info.need(64)) return fail("out of device memory");
        exb_scan_result res;
        double t0 = now();
        if (format == 2) {
            const bool numeric = needs_numeric() || has_computed(EXB_C_GC_CONTENT) || has_computed(EXB_C_MEAN_QUALITY);
            int64_t rec_cap = n / 32 + 4096;
            int64_t info[2] = {0, 0};
            for (int attempt = 0;; attempt++) {
                NvtxRange nv("exb:scan");
                if (attempt && !d_ws.need(exb_fastq_workspace_bytes(n + 16, n + 1))) return fail("out of device memory");
                if (!d_line.need(rec_cap * 4 * 4)) return fail("out of device memory");
                if (numeric)
                    for (int k = 0; k < 4; k++)
                        if (!d_arr[k].need(rec_cap * 4)) return fail("out of device memory");
                if (!rc(exb_fastq_scan(d_cur, 0, n, is_final ? 1 : 0, nullptr, ~0ull, EXB_F_LINES | (numeric ? (EXB_F_SEQ | EXB_F_QUAL) : 0),
                                       d_line.p, rec_cap * 4, 0, d_arr[0].as<uint32_t>(), d_arr[1].as<uint32_t>(), d_arr[2].as<uint32_t>(),
                                       d_arr[3].as<int32_t>(), rec_cap, d_ws.p, d_ws.cap, st)))
                    return false;
                // where the last complete record ends: fetched together with the result block (one sync)
                if (!cu(fastq_chunk_info_launch(d_ws.p, d_line.as<uint32_t>(), d_info.as<int64_t>(), st), "chunk_info")) return false;
                if (!queue_small(0, d_ws.p, (int)(sizeof(exb_scan_result) / 8)) || !queue_small(24, d_info.p, 2) || !z_queue_status() || !wait_small())
                    return false;
                if (!z_check(fname)) return false;  // a corrupt gzip member: its text is garbage, report that and not what the parser made of it
                memcpy(&res, small(0), sizeof(res));
                res.err_pos = ~res.err_pos;  // the device keeps it inverted (see ScanResult)
                info[0] = small(24)[0];
                info[1] = small(24)[1];
                if (!res.overflow) break;
                if (attempt) return fail("internal: record capacity");
                rec_cap = n / 4 + 16;
            }
            t_scan += now() - t0;
            mark(2);
            mark(3);
            mark(4);
            if (res.err_pos != ~0ull)
                return fail("invalid FASTQ record at byte " + std::to_string(cur_file_pos + (int64_t)res.err_pos) + " of " + fname + shard_note());
            int64_t R = (int64_t)(res.total_lines / 4);
            if (is_final) {
                if (res.total_lines % 4 != 0) return fail("unexpected EOF in FASTQ record of " + fname + shard_note());
                consumed = n;
            } else {
                if (R == 0) { grew = true; consumed = 0; return true; }
                consumed = info[0];
            }
            if (count_only.load() && root < 0) {  // COUNT(*) without a filter: the scan's record count is the answer
                item->counted = R;
                return true;
            }
            t0 = now();
            if (!d_lens.need(std::max<int64_t>(R, 1) * 16) || !d_starts.need(std::max<int64_t>(R, 1) * 32) || !d_valid.need(std::max<int64_t>(R, 1)))
                return fail("out of device memory");
            if (!rc(exb_fastq_fields(d_cur, 0, n, d_line.p, 0, nullptr, R, d_lens.as<uint32_t>(), d_valid.as<uint8_t>(), d_starts.as<int64_t>(), d_ws.p, st)))
                return false;
            const uint8_t* bufs[4] = {d_cur, d_cur, d_cur, d_cur};
            const int64_t* o_st; const uint32_t* o_ln; const uint8_t* o_val; const int64_t* o_sel; int64_t o_n;
            if (!select(bufs, R, o_st, o_ln, o_val, o_sel, o_n)) return false;
            t_select += now() - t0;
            mark(3);
            t0 = now();
            const bool ok = materialise(bufs, o_st, o_ln, o_val, o_sel, o_n, item);
            t_mat += now() - t0;
            mark(4);
            return ok;
        }
        // ---- FASTA
        // the sequence column is compacted by the scan itself; skip that when nobody reads the bytes
        // (projection push-down, COUNT(*); string predicates on `sequence` and mapped copies of it still need them)
        bool want_seq = ((column_mask >> 2) & 1u) != 0 && !count_only.load();
        for (const Node& nd : nodes)
            if (nd.kind == N_STR && nd.col == 2) want_seq = true;
        if (has_computed(EXB_C_SEQ_MAP) && !count_only.load()) want_seq = true;
        int64_t rec_cap = n / 64 + 4096;
        for (int attempt = 0;; attempt++) {
            NvtxRange nv("exb:scan");
            if (!d_hdr_start.need(rec_cap * 8) || !d_hdr_end.need(rec_cap * 8) || !d_seq_off.need((rec_cap + 1) * 8) ||
                !d_gc_prefix.need((rec_cap + 1) * 8) || (want_seq && !d_seq.need(n + 64)))
                return fail("out of device memory");
            if (!rc(exb_fasta_scan(d_cur, 0, n, is_final ? 1 : 0, n, nullptr, d_hdr_start.as<int64_t>(), d_hdr_end.as<int64_t>(),
                                   d_seq_off.as<int64_t>(), d_gc_prefix.as<int64_t>(), rec_cap, want_seq ? d_seq.as<uint8_t>() : nullptr,
                                   want_seq ? n + 64 : 0, d_ws.p, d_ws.cap, st)))
                return false;
            if (!queue_small(0, d_ws.p, (int)(sizeof(exb_scan_result) / 8)) || !z_queue_status() || !wait_small()) return false;
            if (!z_check(fname)) return false;
            memcpy(&res, small(0), sizeof(res));
            res.err_pos = ~res.err_pos;  // the device keeps it inverted (see ScanResult)
            if (!res.overflow) break;
            if (attempt) return fail("internal: record capacity");
            rec_cap = n / 2 + 16;
        }
        t_scan += now() - t0;
        mark(2);
        mark(3);
        mark(4);
        if (res.err_pos != ~0ull)
            return fail("invalid FASTA input (missing '>' prefix) at byte " + std::to_string(cur_file_pos + (int64_t)res.err_pos) + " of " + fname);
        int64_t R = (int64_t)res.n_records;
        if (is_final) consumed = n;
        else {
            if (R <= 1) { grew = true; consumed = 0; return true; }
            R -= 1;  // the last record may continue in the next chunk
            if (!queue_small(24, d_hdr_start.as<int64_t>() + R, 1) || !wait_small()) return false;
            consumed = small(24)[0];
        }
        t0 = now();
        if (!d_lens.need(std::max<int64_t>(R, 1) * 12) || !d_starts.need(std::max<int64_t>(R, 1) * 24) || !d_valid.need(std::max<int64_t>(R, 1)) ||
            !d_err.need(8))
            return fail("out of device memory");
        if (!rc(exb_fasta_headers(d_cur, n, d_hdr_start.as<int64_t>(), d_hdr_end.as<int64_t>(), R, d_lens.as<uint32_t>(), d_starts.as<int64_t>(),
                                  d_valid.as<uint8_t>(), d_err.as<uint64_t>(), st)))
            return false;
        if (!cu(fasta_seq_ranges_launch(d_seq_off.as<int64_t>(), nullptr, R, d_starts.as<int64_t>() + 2 * R, d_lens.as<uint32_t>() + 2 * R, st),
                "fasta_seq_ranges"))
            return false;
        if (!queue_small(26, d_err.p, 1) || !wait_small()) return false;
        const uint64_t bad = (uint64_t)small(26)[0];
        if (bad != ~0ull) return fail("FASTA definition without a name at byte " + std::to_string(cur_file_pos + (int64_t)bad) + " of " + fname);
        const uint8_t* bufs[3] = {d_cur, d_cur, d_seq.as<uint8_t>()};
        const int64_t* o_st; const uint32_t* o_ln; const uint8_t* o_val; const int64_t* o_sel; int64_t o_n;
        if (!select(bufs, R, o_st, o_ln, o_val, o_sel, o_n)) return false;
        t_select += now() - t0;
        t0 = now();
        const bool ok = materialise(bufs, o_st, o_ln, o_val, o_sel, o_n, item);
        t_mat += now() - t0;
        return ok;
    }
    std::string shard_note() const {
        if (range_lo <= 0 && range_hi <= 0) return "";
        return " (byte-range shard [" + std::to_string(shard_begin) + ", " + std::to_string(shard_end) +
               "): if the file is well formed, the shard cut could not be resynchronised -- rerun with one reader)";
    }

    // A chunk = the unconsumed tail of the previous chunk (it never left HBM: a device-to-device move to the front of
    // the other input buffer) followed by the next raw block.  While chunk k is scanned, split and copied back, block
    // k+1 -- if the IO thread already has it -- crosses PCIe on the copy stream into a staging buffer, so H2D of k+1
    // overlaps the kernels of k and the D2H of k-1 (two copy engines, opposite directions).
    struct Active {  // (page-cache registration waits while any reader thread runs: FileMap::start_register)
        Active() { FileMap::io_active().fetch_add(1); }
        ~Active() { FileMap::io_active().fetch_sub(1); }
    };
    void dev_main() {
        Active active;
        std::shared_ptr<HostBlock> host_cur;  // pinned block whose bytes may still be in flight to the device (chunk results that
                                              // borrow its bytes hold it too: see host_base)
        Block staged;                 // the prefetched block (its pinned buffer included); valid when have_staged
        bool have_staged = false, staged_on_device = false;
        int cur = 0;
        int64_t carry_off = 0, carry_len = 0;  // the tail lives in d_inb[cur] at [carry_off, carry_off + carry_len)
        auto finish = [&](const std::string& err_in) {
            std::string err = err_in;
            if (err.empty() && z_pending && !(z_queue_status() && wait_small() && z_check(files[std::min(cur_file, files.size() - 1)]))) err = derr;
            if (err.empty() && !flush_pending(0)) err = derr;
            cudaStreamSynchronize(sc);
            cudaStreamSynchronize(st);
            cudaStreamSynchronize(sd);
            host_cur.reset();
            host_block.reset();
            host_base = nullptr;
            if (have_staged) pool->put(staged.h);
            pend.clear();
            free_device();
            OutItem e;
            e.end = true;
            e.error = err;
            outq.push(std::move(e));
        };
        {
            const double ti = now();
            const bool ok = init_device();
            t_init = now() - ti;
            if (!ok) return finish(derr);
        }
        while (!stopping) {
            Block b;
            bool on_device = false;
            if (have_staged) {
                b = std::move(staged);
                on_device = staged_on_device;
                have_staged = false;
            } else {
                if (!flush_pending(0)) return finish(derr);  // nothing to overlap with: do not sit on finished results
                double t0 = now();
                const bool popped = inq.pop(b);
                t_dev_pop += now() - t0;
                if (!popped) break;
            }
            if (b.end) return finish(b.error);
            double t0 = now();
            NvtxRange nv("exb:chunk");
            // ---- assemble chunk k in the other input buffer
            mark(0);
            const int nxt = cur ^ 1;
            const int64_t n = carry_len + b.raw_len;
            if (!d_inb[nxt].need(n + 64)) {
                pool->put(b.h);
                return finish("out of device memory");
            }
            uint8_t* dst = d_inb[nxt].as<uint8_t>();
            bool ok = true;
            if (carry_len) ok = cu(cudaMemcpyAsync(dst, d_inb[cur].as<uint8_t>() + carry_off, (size_t)carry_len, cudaMemcpyDeviceToDevice, st), "D2D tail");
            if (ok && b.zlen) {  // BGZF: the members are inflated straight into the chunk
                if (on_device) ok = cu(cudaStreamWaitEvent(st, ev_staged, 0), "wait") && z_inflate(b, d_zin[1], dst + carry_len, st) && cu(cudaEventRecord(ev_stage_free, st), "record");
                else ok = z_copy_in(b, d_zin[0], st) && z_inflate(b, d_zin[0], dst + carry_len, st);
            } else if (ok && b.raw_len) {
                if (on_device) {
                    ok = cu(cudaStreamWaitEvent(st, ev_staged, 0), "wait") &&
                         cu(cudaMemcpyAsync(dst + carry_len, d_stage.p, (size_t)b.raw_len, cudaMemcpyDeviceToDevice, st), "D2D block") &&
                         cu(cudaEventRecord(ev_stage_free, st), "record");
                } else {
                    ok = cu(FileMap::host_to_device(dst + carry_len, b.data(), b.raw_len, b.ext ? b.ext_map.get() : nullptr, st), "H2D");
                }
            }
            // ---- the same chunk on the host: the block's bytes are already there; the carried tail (at most the record that
            // straddles the two blocks) is copied in front of them, into the block's slack.  Not possible when the tail is
            // larger than the slack: that chunk's strings are gathered and copied back as before.
            {
                uint8_t* blk = b.data();
                uint8_t* base = nullptr;
                if (b.zlen) {  // the host holds the compressed bytes only: strings are gathered on the device and copied back
                } else if (carry_len == 0) {
                    base = blk;
                } else if (b.ext) {  // the file is one contiguous piece of host memory: the carried tail sits right in front
                    if (b.ext - b.ext_map->p >= carry_len) base = blk - carry_len;
                } else if (carry_len <= b.data_off && host_base) {
                    memcpy(blk - carry_len, host_base + carry_off, (size_t)carry_len);
                    base = blk - carry_len;
                }
                host_base = base;
            }
            host_cur = std::make_shared<HostBlock>(b.h, pool, b.ext_map);  // (the previous chunk synchronised `st` after its copies: that block
            host_block = host_cur;                              //  goes back to the pool once no chunk result points into it)
            if (!ok) return finish(derr);
            mark(1);
            cur = nxt;
            d_cur = dst;
            cur_file = b.file_idx;
            cur_file_pos = b.raw_file_pos - carry_len;
            carry_off = carry_len = 0;
            // ---- prefetch block k+1 while chunk k is processed
            if (!b.eof || b.file_idx + 1 < files.size()) {
                Block nb;
                if (inq.try_pop(nb)) {
                    staged_on_device = false;
                    if (!nb.end && nb.zlen > 0) {
                        if (cu(cudaStreamWaitEvent(sc, ev_stage_free, 0), "wait") && z_copy_in(nb, d_zin[1], sc) && cu(cudaEventRecord(ev_staged, sc), "record"))
                            staged_on_device = true;
                        else
                            return finish(derr);
                    } else if (!nb.end && nb.raw_len > 0 && d_stage.need(nb.raw_len + 64)) {
                        // the staging buffer is free once the previous staged block has been moved out of it
                        if (cu(cudaStreamWaitEvent(sc, ev_stage_free, 0), "wait") &&
                            cu(FileMap::host_to_device(d_stage.p, nb.data(), nb.raw_len, nb.ext ? nb.ext_map.get() : nullptr, sc), "H2D prefetch") &&
                            cu(cudaEventRecord(ev_staged, sc), "record"))
                            staged_on_device = true;
                        else
                            return finish(derr);
                    }
                    staged = std::move(nb);
                    have_staged = true;
                }
            }
            if (n == 0) continue;  // empty file
            int64_t consumed = 0;
            bool grew = false;
            OutItem item;
            if (!process_chunk(n, b.eof, consumed, grew, &item)) return finish(derr);
            if (grew) {  // no complete record in the chunk: keep all of it and read bigger blocks from now on
                consumed = 0;
                block_bytes.store(std::min<int64_t>(block_bytes.load() * 2, 1ll << 30));
            }
            carry_off = consumed;
            carry_len = n - consumed;
            bytes_done.fetch_add(b.zlen ? b.zlen : consumed);  // (progress is in FILE bytes: a block of BGZF members counts as its compressed size)
            t_dev_work += now() - t0;
            if (trace >= 2) {
                mark(5);
                cudaEventSynchronize(tev[5]);
                float a = 0, b2 = 0, c = 0, d = 0, e2 = 0;
                cudaEventElapsedTime(&a, tev[0], tev[1]);
                cudaEventElapsedTime(&b2, tev[1], tev[2]);
                cudaEventElapsedTime(&c, tev[2], tev[3]);
                cudaEventElapsedTime(&d, tev[3], tev[4]);
                cudaEventElapsedTime(&e2, tev[4], tev[5]);
                fprintf(stderr, "exon_b200 chunk %lld bytes (%s): host %.3f ms | gpu: copy-in %.3f scan %.3f fields+select %.3f materialise %.3f tail %.3f ms\n",
                        (long long)n, on_device ? "prefetched" : "direct H2D", (now() - t0) * 1e3, a, b2, c, d, e2);
            }
            if (item.counted) {  // (materialised results travel through `pend`)
                t0 = now();
                const bool pushed = outq.push(std::move(item));
                t_dev_push += now() - t0;
                if (!pushed) break;
            }
        }
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(sd);
        host_cur.reset();
        host_block.reset();
        host_base = nullptr;
        if (have_staged) pool->put(staged.h);
        pend.clear();
        free_device();
    }

    // ---- COUNT(*) [WHERE predicates on the quality line]: the C2 query.  Nothing per record is needed, so the chunks are
    // not re-cut at record boundaries at all: block k goes to the device and is scanned as the next RANGE of a chained
    // scan (exb_fastq_scan_filter: the fused byte pass judges every quality line as it is emitted; the previous range's
    // 128-byte result block carries the open line and the line count across the edge), the aggregates accumulate on the
    // device, and the host looks at them once per file.  No carry copy, no field split, no selection, no per-chunk sync:
    // the device thread costs one H2D copy + one ~40 us scan per 64 MiB, so the pipeline runs at the rate of its slowest
    // copy (page cache -> pinned, or PCIe).
    bool fused_count_plan(std::vector<exb_predicate>* out) const {
        if (format != 2 || !count_only.load()) return false;
        if (getenv("EXON_B200_NO_FUSED_COUNT")) return false;
        out->clear();
        if (root < 0) return true;
        // AND-only tree of numeric leaves on the quality line
        std::vector<int> stack{root};
        while (!stack.empty()) {
            const Node& nd = nodes[stack.back()];
            stack.pop_back();
            if (nd.kind == N_AND) {
                stack.push_back(nd.lhs);
                stack.push_back(nd.rhs);
            } else if (nd.kind == N_NUM && (nd.field == EXB_P_MEAN_QUALITY || nd.field == EXB_P_QUAL_LEN)) {
                out->push_back(exb_predicate{nd.field, nd.op, nd.value});
            } else {
                return false;
            }
        }
        return out->size() <= EXB_MAX_PREDICATES;
    }
    void dev_main_fused(std::vector<exb_predicate> preds) {
        Active active;
        DBuf d_chunk[2], d_wsx[2], d_agg;
        struct InFlight {
            HBuf* h;                       // pinned block (nullptr: the bytes came straight from a registered file mapping,
            std::shared_ptr<FileMap> map;  //  which this keeps alive until the copy is done)
            cudaEvent_t ev;
        };
        cudaEvent_t ev_zc = nullptr, ev_zf[2] = {nullptr, nullptr};  // BGZF: members copied / staging buffer free again
        std::deque<InFlight> inflight;  // pinned blocks whose H2D copy may still be running
        std::vector<cudaEvent_t> ev_pool;
        auto release_done = [&](size_t keep) {
            while (inflight.size() > keep) {
                cudaEventSynchronize(inflight.front().ev);
                pool->put(inflight.front().h);
                ev_pool.push_back(inflight.front().ev);
                inflight.pop_front();
            }
        };
        auto finish = [&](const std::string& err, int64_t count) {
            cudaStreamSynchronize(st);
            release_done(0);
            for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
            cudaStreamSynchronize(sc);
            for (cudaEvent_t e : {ev_zc, ev_zf[0], ev_zf[1]})
                if (e) cudaEventDestroy(e);
            for (DBuf* b : {&d_chunk[0], &d_chunk[1], &d_wsx[0], &d_wsx[1], &d_agg}) b->release();
            free_device();
            if (err.empty() && count > 0) {
                OutItem c;
                c.counted = count;
                outq.push(std::move(c));
            }
            OutItem e;
            e.end = true;
            e.error = err;
            outq.push(std::move(e));
        };
        {
            const double ti = now();
            const bool ok = init_device();
            t_init = now() - ti;
            if (!ok) return finish(derr, 0);
        }
        if (!d_agg.need(64) || !cu(cudaMemsetAsync(d_agg.p, 0, 64, st), "memset")) return finish(derr.empty() ? "out of device memory" : derr, 0);
        int kb = 0, kw = 0;              // chunk buffers alternate every block, scan workspaces every scan
        const void* prev_ws = nullptr;   // workspace of the previous range of this file's chain
        const uint8_t* prev_data = nullptr;  // previous chunk buffer: prev_data[0] is file offset prev_pos
        int64_t prev_pos = 0;
        int64_t tail = 0;  // bytes at the end of the previous buffer that were not scanned yet: a block of inflated BGZF members
                           // ends anywhere, a chained range must end on a multiple of 16, so the last < 16 bytes wait for the next block
        while (!stopping) {
            Block b;
            double t0 = now();
            const bool popped = inq.pop(b);
            t_dev_pop += now() - t0;
            if (!popped) break;
            if (b.end) {
                if (!b.error.empty()) return finish(b.error, 0);
                break;
            }
            if (b.raw_len == 0 && !prev_ws && tail == 0 && !b.zlen) {  // empty file
                pool->put(b.h);
                continue;
            }
            t0 = now();
            NvtxRange nv("exb:count_chunk");
            const int i = kb & 1;
            const int64_t pos = b.raw_file_pos - tail, n = tail + b.raw_len;
            if (!d_chunk[i].need(n + 128)) {
                pool->put(b.h);
                return finish("out of device memory", 0);
            }
            // data at a 16-aligned address + (pos % 16), so that the buffer viewed from file offset 0 is 16-byte aligned;
            // the 16 bytes in front of it are the end of the previous range (the scan looks one byte back for CR LF)
            uint8_t* data = d_chunk[i].as<uint8_t>() + 32 + (pos & 15);
            bool ok = true;
            if (prev_data && (prev_ws || tail)) {  // halo + the unscanned tail, in one piece
                const int64_t h = prev_ws ? 16 : 0;  // (a chain has scanned >= 16 bytes; they are in the previous buffer or in ITS halo)
                if (h + tail > 0)
                    ok = cu(cudaMemcpyAsync(data - h, prev_data + (pos - prev_pos) - h, (size_t)(h + tail), cudaMemcpyDeviceToDevice, st), "D2D halo");
            }
            if (ok && b.zlen) {
                // the compressed members cross PCIe on the copy stream while the compute stream still inflates the block before:
                // two staging buffers, an event per buffer says when its last reader (the inflate kernel) is done
                const int z = kb & 1;
                if (!ev_zc) {
                    ok = cu(cudaEventCreateWithFlags(&ev_zc, cudaEventDisableTiming), "event") && cu(cudaEventCreateWithFlags(&ev_zf[0], cudaEventDisableTiming), "event") &&
                         cu(cudaEventCreateWithFlags(&ev_zf[1], cudaEventDisableTiming), "event");
                }
                ok = ok && cu(cudaStreamWaitEvent(sc, ev_zf[z], 0), "wait") && z_copy_in(b, d_zin[z], sc) && cu(cudaEventRecord(ev_zc, sc), "record") &&
                     cu(cudaStreamWaitEvent(st, ev_zc, 0), "wait") && z_inflate(b, d_zin[z], data + tail, st) && cu(cudaEventRecord(ev_zf[z], st), "record");
            }
            else if (ok && b.raw_len) ok = cu(FileMap::host_to_device(data + tail, b.data(), b.raw_len, b.ext ? b.ext_map.get() : nullptr, st), "H2D");
            cudaEvent_t ev = nullptr;
            if (!ev_pool.empty()) {
                ev = ev_pool.back();
                ev_pool.pop_back();
            } else if (ok) {
                ok = cu(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
            }
            if (ok) ok = cu(cudaEventRecord(ev, st), "record");
            if (!ok) {
                pool->put(b.h);
                return finish(derr, 0);
            }
            inflight.push_back(InFlight{b.h, b.ext_map, ev});
            const int64_t scan_end = (b.eof || !b.zlen) ? pos + n : ((pos + n) & ~(int64_t)15);
            if (scan_end > pos || b.eof) {
                const int w = kw & 1;
                if (!d_wsx[w].need(exb_scan_workspace_bytes(scan_end - pos + 64))) return finish("out of device memory", 0);
                const uint8_t* base = data - pos;  // base[file offset] = that byte
                if (!rc(exb_fastq_scan_filter(base, pos, scan_end, b.eof ? 1 : 0, prev_ws, preds.empty() ? nullptr : preds.data(), (int)preds.size(),
                                              d_agg.as<int64_t>(), 1, d_wsx[w].p, d_wsx[w].cap, st)))
                    return finish(derr, 0);
                prev_ws = d_wsx[w].p;
                kw++;
            }
            tail = pos + n - std::max(scan_end, pos);
            prev_data = data;
            prev_pos = pos;
            bytes_done.fetch_add(b.zlen ? b.zlen : b.raw_len);
            release_done(2);
            kb++;
            if (b.eof) {  // the file (or shard) ends here: its line count and error marks, once per file
                exb_scan_result res;
                if (!queue_small(0, prev_ws, (int)(sizeof(exb_scan_result) / 8)) || !z_queue_status() || !wait_small()) return finish(derr, 0);
                const std::string& fname = files[b.file_idx];
                if (!z_check(fname)) return finish(derr, 0);
                memcpy(&res, small(0), sizeof(res));
                res.err_pos = ~res.err_pos;
                if (res.err_pos != ~0ull) return finish("invalid FASTQ record in " + fname + shard_note(), 0);
                if (res.total_lines % 4 != 0) return finish("unexpected EOF in FASTQ record of " + fname + shard_note(), 0);
                prev_ws = nullptr;  // the next file starts a new chain
                prev_data = nullptr;
                tail = 0;
            }
            t_dev_work += now() - t0;
        }
        if (stopping) return finish("", 0);
        if (!queue_small(16, d_agg.p, 8) || !wait_small()) return finish(derr, 0);
        finish("", small(16)[0]);
    }

    // ------------------------------------------------------------------ caller (under call_mu)
    // make rows available; false = end of stream or error (check `error`)
    bool advance() {
        if (finished) return false;
        if (!started) {
            if (!check_device()) {
                finished = true;
                return false;
            }
            started = true;
            block_bytes.store(std::max<int64_t>(chunk_bytes & ~(int64_t)15, 4096));
            io_thread = std::thread([this] { io_main(); });
            std::vector<exb_predicate> fused_preds;
            if (fused_count_plan(&fused_preds)) dev_thread = std::thread([this, fused_preds] { dev_main_fused(fused_preds); });
            else dev_thread = std::thread([this] { dev_main(); });
        }
        while (next_row >= rows) {
            rows = next_row = 0;
            cur.reset();
            OutItem it;
            const double t0 = now();
            const bool popped = outq.pop(it);
            t_call_pop += now() - t0;
            if (!popped) {
                finished = true;
                return false;
            }
            if (it.end) {
                finished = true;
                error = it.error;
                return false;
            }
            counted += it.counted;
            rows_out += it.counted;
            if (it.res) {
                cur = it.res;
                rows = cur->rows;
                rows_out += rows;
            }
        }
        return true;
    }
};

// ------------------------------------------------------------------ Arrow export
struct BatchHolder {
    int ncols = 0;
    std::vector<int32_t> off[4];
    std::vector<uint8_t> data[4];
    std::vector<uint8_t> validity;  // bitmap of the description column
    ArrowArray child[4];
    ArrowArray* child_ptr[4];
    const void* child_bufs[4][3];
    const void* root_bufs[1];
};
void release_child(ArrowArray* a) { a->release = nullptr; }
void release_batch(ArrowArray* a) {
    BatchHolder* h = reinterpret_cast<BatchHolder*>(a->private_data);
    for (int c = 0; c < h->ncols; c++)
        if (h->child[c].release) h->child[c].release(&h->child[c]);
    delete h;
    a->release = nullptr;
}

struct SchemaHolder {
    std::vector<std::string> names;
    ArrowSchema child[4];
    ArrowSchema* child_ptr[4];
};
void release_schema_child(ArrowSchema* s) { s->release = nullptr; }
void release_schema(ArrowSchema* s) {
    SchemaHolder* h = reinterpret_cast<SchemaHolder*>(s->private_data);
    delete h;
    s->release = nullptr;
}

int stream_get_schema(ArrowArrayStream* s, ArrowSchema* out) {
    Reader* r = reinterpret_cast<Reader*>(s->private_data);
    SchemaHolder* h = new SchemaHolder();
    h->names = r->col_names;
    for (int c = 0; c < r->ncols; c++) {
        ArrowSchema& k = h->child[c];
        memset(&k, 0, sizeof(k));
        k.format = "u";
        k.name = h->names[c].c_str();
        k.flags = 2;  // ARROW_FLAG_NULLABLE
        k.release = release_schema_child;
        h->child_ptr[c] = &k;
    }
    memset(out, 0, sizeof(*out));
    out->format = "+s";
    out->name = "";
    out->n_children = r->ncols;
    out->children = h->child_ptr;
    out->release = release_schema;
    out->private_data = h;
    return 0;
}

int stream_get_next(ArrowArrayStream* s, ArrowArray* out) {
    Reader* r = reinterpret_cast<Reader*>(s->private_data);
    memset(out, 0, sizeof(*out));
    std::lock_guard<std::mutex> guard(r->call_mu);
    if (!r->advance()) {
        if (!r->error.empty()) return 5;  // EIO; message through get_last_error
        out->release = nullptr;           // end of stream
        return 0;
    }
    // rows [next_row, next_row + k): at most batch_size rows and < 2 GiB per column (utf8 has int32 offsets)
    int64_t b = r->next_row, e = std::min(r->rows, b + r->batch_size);
    for (int c = 0; c < r->ncols; c++) {
        const int64_t* off = r->cur->cols[c].off;
        if (!off) continue;
        while (e > b + 1 && off[e] - off[b] > 0x7FFFFFF0ll) e = b + (e - b) / 2;
        if (off[e] - off[b] > 0x7FFFFFF0ll) {
            r->error = "a single " + r->col_names[c] + " value exceeds the 2 GiB limit of Arrow utf8";
            return 5;
        }
    }
    const int64_t k = e - b;
    BatchHolder* h = new BatchHolder();
    h->ncols = r->ncols;
    int64_t nulls = 0;
    h->validity.assign((size_t)((k + 7) / 8), 0);
    const uint8_t* desc_valid = nullptr;
    for (int c = 0; c < r->ncols; c++)
        if (r->cur->cols[c].valid) desc_valid = r->cur->cols[c].valid;
    for (int64_t i = 0; i < k; i++) {
        if (!desc_valid || desc_valid[b + i]) h->validity[i >> 3] |= (uint8_t)(1u << (i & 7));
        else nulls++;
    }
    for (int c = 0; c < r->ncols; c++) {
        const OutCol& col = r->cur->cols[c];
        h->off[c].assign((size_t)k + 1, 0);
        if (col.off) {
            const int64_t base = col.off[b];
            for (int64_t i = 0; i <= k; i++) h->off[c][i] = (int32_t)(col.off[b + i] - base);
            h->data[c].assign(col.data + base, col.data + col.off[e]);
        }
        const bool is_desc = r->col_names[c] == "description";
        ArrowArray& a = h->child[c];
        memset(&a, 0, sizeof(a));
        a.length = k;
        a.null_count = is_desc ? nulls : 0;
        a.n_buffers = 3;
        h->child_bufs[c][0] = (is_desc && nulls > 0) ? h->validity.data() : nullptr;
        h->child_bufs[c][1] = h->off[c].data();
        h->child_bufs[c][2] = h->data[c].empty() ? (const void*)"" : (const void*)h->data[c].data();
        a.buffers = h->child_bufs[c];
        a.release = release_child;
        h->child_ptr[c] = &a;
    }
    out->length = k;
    out->null_count = 0;
    out->n_buffers = 1;
    h->root_bufs[0] = nullptr;
    out->buffers = h->root_bufs;
    out->n_children = r->ncols;
    out->children = h->child_ptr;
    out->release = release_batch;
    out->private_data = h;
    r->next_row = e;
    return 0;
}
const char* stream_last_error(ArrowArrayStream* s) {
    Reader* r = reinterpret_cast<Reader*>(s->private_data);
    return r->error.empty() ? nullptr : r->error.c_str();
}
void stream_release(ArrowArrayStream* s) {
    delete reinterpret_cast<Reader*>(s->private_data);
    s->release = nullptr;
}

ReaderResult reader_error(const std::string& m) {
    ReaderResult r;
    r.error = dup_cstr(m);
    return r;
}

}  // namespace

extern "C" {

void exb_free_string(const char* s) { free((void*)s); }

// arrow_reader.rs:173-197: take the text after the last '.', if it names a
// compression codec step to the previous '.'-separated component, map that to a
// file type; the result is the upper-case type name or NULL.
ReplacementScanResult replacement_scan(const char* uri) {
    ReplacementScanResult r;
    r.file_type = nullptr;
    if (!uri) return r;
    std::string u(uri);
    size_t dot = u.rfind('.');
    std::string ext = dot == std::string::npos ? u : u.substr(dot + 1);
    std::string rest = dot == std::string::npos ? std::string() : u.substr(0, dot);
    if (compression_from_str(ext) != 0) {
        size_t dot2 = rest.rfind('.');
        ext = dot2 == std::string::npos ? rest : rest.substr(dot2 + 1);
    }
    int ft = file_type_from_str(ext);
    if (ft == 1) r.file_type = dup_cstr("FASTA");
    if (ft == 2) r.file_type = dup_cstr("FASTQ");
    return r;
}

// Shared by new_reader and exb_reader_open: resolves compression / format, lists the files, parses the filter.
static Reader* open_reader(const char* uri, uintptr_t batch_size, const char* compression, const char* file_format, const char* filters,
                           std::string* err) {
    if (!uri || !file_format) {
        *err = "new_reader: null argument";
        return nullptr;
    }
    std::string u(uri);
    // arrow_reader.rs:60-91: NULL compression = infer from the text after the last '.'
    int comp;
    if (!compression) {
        size_t dot = u.rfind('.');
        std::string ext = dot == std::string::npos ? u : u.substr(dot + 1);
        comp = ext == "gz" ? 1 : (ext == "zst" ? 2 : 0);
    } else {
        comp = compression_from_str(compression);
    }
    const int ft = file_type_from_str(file_format);
    if (ft == 0) {
        *err = std::string("could not parse file_format ") + file_format;
        return nullptr;
    }

    std::unique_ptr<Reader> r(new Reader());
    r->format = ft;
    r->batch_size = batch_size > 0 ? (int64_t)batch_size : 2048;
    if (ft == 1) r->col_names = {"id", "description", "sequence"};
    else r->col_names = {"name", "description", "sequence", "quality_scores"};
    r->ncols = (int)r->col_names.size();

    struct stat sb;
    if (u.empty() || stat(u.c_str(), &sb) != 0) {
        *err = "could not register table: no such file or directory: " + u;
        return nullptr;
    }
    if (S_ISDIR(sb.st_mode)) {
        // listing table: every file whose name carries the format's extension (+ the codec's)
        std::vector<std::string> exts = ft == 1 ? std::vector<std::string>{".fasta", ".fa", ".fna"} : std::vector<std::string>{".fastq", ".fq"};
        const char* csuf = comp == 1 ? ".gz" : (comp == 2 ? ".zst" : (comp == 3 ? ".bz2" : (comp == 4 ? ".xz" : "")));
        DIR* d = opendir(u.c_str());
        if (!d) {
            *err = "could not list " + u;
            return nullptr;
        }
        std::vector<std::string> names;
        while (dirent* de = readdir(d)) {
            std::string nm = de->d_name;
            for (const std::string& x : exts)
                if (ends_with(lower(nm), x + csuf)) {
                    names.push_back(nm);
                    break;
                }
        }
        closedir(d);
        std::sort(names.begin(), names.end());
        for (const std::string& nm : names) {
            r->files.push_back(u + (ends_with(u, "/") ? "" : "/") + nm);
            r->file_comp.push_back(comp);
        }
    } else {
        r->files.push_back(u);
        r->file_comp.push_back(comp);
    }

    if (filters && filters[0]) {
        std::string f(filters);
        Parser p(f, r->nodes, r->col_names);
        r->root = p.parse_or();
        p.ws();
        if (r->root < 0 || p.i != f.size()) {
            *err = "could not execute sql: cannot parse filter `" + f + "`: " + (p.err.empty() ? "trailing text" : p.err);
            return nullptr;
        }
    }
    const char* cb = getenv("EXON_B200_CHUNK_BYTES");
    if (cb && atoll(cb) > 0) r->chunk_bytes = atoll(cb);
    return r.release();
}

ReaderResult new_reader(ArrowArrayStream* stream_ptr, const char* uri, uintptr_t batch_size, const char* compression,
                        const char* file_format, const char* filters) {
    if (!stream_ptr) return reader_error("new_reader: null argument");
    std::string err;
    Reader* r = open_reader(uri, batch_size, compression, file_format, filters, &err);
    if (!r) return reader_error(err);
    stream_ptr->get_schema = stream_get_schema;
    stream_ptr->get_next = stream_get_next;
    stream_ptr->get_last_error = stream_last_error;
    stream_ptr->release = stream_release;
    stream_ptr->private_data = r;
    ReaderResult ok;
    ok.error = nullptr;
    return ok;
}

// ---- the same reader without the Arrow detour (include/exon_b200.h, "native reader")
struct exb_reader {
    Reader* r;
    std::string err;
};

static int apply_options(Reader* r, const exb_reader_options* o, std::string* err) {
    if (!o) return 0;
    if (o->size < sizeof(exb_reader_options)) {
        *err = "exb_reader_open2: options struct is older than this library";
        return -1;
    }
    r->device = o->device;
    r->column_mask = o->column_mask;
    r->flags = o->flags;
    if (o->n_computed < 0 || o->n_computed > EXB_MAX_COMPUTED) {
        *err = "exb_reader_open2: too many computed columns";
        return -1;
    }
    for (int i = 0; i < o->n_computed; i++) {
        const exb_computed& c = o->computed[i];
        const bool fastq_only = c.kind == EXB_C_QUALITY_LIST || c.kind == EXB_C_MEAN_QUALITY || c.kind == EXB_C_QUAL_LENGTH;
        if (c.kind < EXB_C_GC_CONTENT || c.kind > EXB_C_QUAL_LENGTH || (fastq_only && r->format != 2) ||
            (c.kind == EXB_C_SEQ_MAP && (c.arg < EXB_MAP_REVERSE_COMPLEMENT || c.arg > EXB_MAP_REVERSE_TRANSCRIBE))) {
            *err = "exb_reader_open2: computed column " + std::to_string(i) + " is not applicable to this file format";
            return -1;
        }
        r->computed.push_back(c);
    }
    if (o->range_lo > 0 || o->range_hi > 0) {
        if (r->files.size() != 1 || r->file_comp[0] != 0) {
            *err = "exb_reader_open2: byte-range shards need one uncompressed file";
            return -1;
        }
        if (o->range_lo < 0 || (o->range_hi > 0 && o->range_hi < o->range_lo)) {
            *err = "exb_reader_open2: bad byte range";
            return -1;
        }
        r->range_lo = o->range_lo;
        r->range_hi = o->range_hi;
    }
    if (o->file_hi > 0 || o->file_lo > 0) {
        const size_t lo = (size_t)std::max(o->file_lo, 0), hi = o->file_hi > 0 ? std::min((size_t)o->file_hi, r->files.size()) : r->files.size();
        if (lo > hi) {
            *err = "exb_reader_open2: bad file range";
            return -1;
        }
        r->files = std::vector<std::string>(r->files.begin() + lo, r->files.begin() + hi);
        r->file_comp = std::vector<int>(r->file_comp.begin() + lo, r->file_comp.begin() + hi);
    }
    return 0;
}

static void measure_input(Reader* r) {
    int64_t total = 0;
    for (const std::string& f : r->files) {
        struct stat sb;
        if (stat(f.c_str(), &sb) == 0) total += sb.st_size;
    }
    if (r->range_lo > 0 || r->range_hi > 0) {
        const int64_t hi = r->range_hi > 0 ? std::min(r->range_hi, total) : total;
        total = std::max<int64_t>(hi - std::min(r->range_lo, hi), 0);
    }
    r->bytes_total.store(total);
}

int exb_reader_open2(const char* uri, const char* file_format, const char* compression, int64_t batch_rows, const char* filters,
                     const exb_reader_options* options, exb_reader** out) {
    if (!out) return set_err(EXB_ERR_ARG, "exb_reader_open: null out");
    *out = nullptr;
    std::string err;
    Reader* r = open_reader(uri, (uintptr_t)(batch_rows > 0 ? batch_rows : 2048), compression, file_format, filters, &err);
    if (!r) return set_err(err.find("no such file") != std::string::npos || err.find("could not list") != std::string::npos ? EXB_ERR_IO : EXB_ERR_ARG,
                           "%s", err.c_str());
    if (apply_options(r, options, &err) != 0) {
        delete r;
        return set_err(EXB_ERR_ARG, "%s", err.c_str());
    }
    measure_input(r);
    exb_reader* h = new exb_reader();
    h->r = r;
    *out = h;
    return 0;
}

int exb_reader_open(const char* uri, const char* file_format, const char* compression, int64_t batch_rows, const char* filters,
                    uint32_t column_mask, exb_reader** out) {
    exb_reader_options o;
    memset(&o, 0, sizeof(o));
    o.size = sizeof(o);
    o.device = -1;
    o.column_mask = column_mask;
    return exb_reader_open2(uri, file_format, compression, batch_rows, filters, &o, out);
}

int exb_reader_plan(const char* uri, const char* file_format, const char* compression, int64_t* total_bytes, int32_t* n_files,
                    int32_t* range_shardable) {
    std::string err;
    Reader* r = open_reader(uri, 2048, compression, file_format, nullptr, &err);
    if (!r) return set_err(err.find("no such file") != std::string::npos || err.find("could not list") != std::string::npos ? EXB_ERR_IO : EXB_ERR_ARG,
                           "%s", err.c_str());
    measure_input(r);
    if (total_bytes) *total_bytes = r->bytes_total.load();
    if (n_files) *n_files = (int32_t)r->files.size();
    if (range_shardable) {
        struct stat sb;
        *range_shardable = r->files.size() == 1 && r->file_comp[0] == 0 && stat(r->files[0].c_str(), &sb) == 0 && S_ISREG(sb.st_mode) ? 1 : 0;
    }
    delete r;
    return 0;
}

int exb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int exb_reader_columns(const exb_reader* h, const char** names, int cap) {
    if (!h) return 0;
    for (int c = 0; c < h->r->ncols && c < cap; c++) names[c] = h->r->col_names[c].c_str();
    return h->r->ncols;
}

static int reader_failure(exb_reader* h) {
    Reader* r = h->r;
    h->err = r->error;
    const bool fmt = r->error.find("invalid FAST") != std::string::npos || r->error.find("unexpected EOF") != std::string::npos ||
                     r->error.find("without a name") != std::string::npos;
    const bool chr = r->error.find("Invalid character in sequence") != std::string::npos;
    return set_err(chr ? EXB_ERR_INVALID_CHAR : (fmt ? EXB_ERR_FORMAT : (r->error.find("CUDA") != std::string::npos ? EXB_ERR_CUDA : EXB_ERR_IO)),
                   "%s", r->error.c_str());
}

int exb_reader_next(exb_reader* h, exb_batch* out) {
    if (!h || !out) return set_err(EXB_ERR_ARG, "exb_reader_next: null argument");
    Reader* r = h->r;
    memset(out, 0, sizeof(*out));
    std::lock_guard<std::mutex> guard(r->call_mu);
    if (!r->advance()) {
        if (!r->error.empty()) return reader_failure(h);
        return 0;  // end of stream: n_rows = 0
    }
    const int64_t b = r->next_row, e = std::min(r->rows, b + r->batch_size);
    const int64_t k = e - b;
    out->n_rows = k;
    out->n_cols = r->ncols;
    out->n_computed = (int32_t)r->computed.size();
    out->batch_index = r->next_batch_index++;
    const bool word_aligned = (b & 63) == 0;
    for (int c = 0; c < r->n_out(); c++) {
        const OutCol& col = r->cur->cols[c];
        exb_column_view& v = out->cols[c];
        v.type = col.type;
        v.kind = col.kind;
        if (!col.present) continue;  // projected out
        const bool has_nulls = col.valid && col.nulls_word && *col.nulls_word != 0;
        if (col.type == EXB_T_VARCHAR) {
            if (col.off) v.offsets = col.off + b;
            v.data = col.data;
            if (col.str) v.strings = col.str + 16 * b;
        } else if (col.type == EXB_T_INT32_LIST) {
            const int64_t kb = b / r->batch_size;
            v.values = col.values + 4 * col.batch_base[kb];
            v.n_values = col.batch_base[kb + 1] - col.batch_base[kb];
            v.list_entries = col.entries + 2 * b;
            if (col.off) v.offsets = col.off + b;
        } else {
            v.values = col.values + (int64_t)col.elem * b;
            v.n_values = k;
        }
        if (col.valid) {
            v.valid = col.valid + b;
            v.chunk_nulls = col.nulls_word ? (int64_t)*col.nulls_word : -1;
            if (has_nulls && word_aligned) v.valid_bits = col.vbits + (b >> 6);
        }
    }
    out->owner = new std::shared_ptr<ChunkResult>(r->cur);
    r->next_row = e;
    return 0;
}

void exb_batch_release(exb_batch* b) {
    if (!b || !b->owner) return;
    delete reinterpret_cast<std::shared_ptr<ChunkResult>*>(b->owner);
    b->owner = nullptr;
}

int exb_reader_count(exb_reader* h, int64_t* n_rows) {
    if (!h || !n_rows) return set_err(EXB_ERR_ARG, "exb_reader_count: null argument");
    Reader* r = h->r;
    std::lock_guard<std::mutex> guard(r->call_mu);
    r->count_only.store(true);
    // rows of chunks that were materialised before the switch are counted where they sit; advance() only returns
    // true with unread rows in hand, so each iteration consumes them
    while (r->advance()) {
        r->counted += r->rows - r->next_row;
        r->next_row = r->rows;
    }
    if (!r->error.empty()) return reader_failure(h);
    *n_rows = r->counted;
    return 0;
}

int exb_reader_progress(const exb_reader* h, int64_t* bytes_done, int64_t* bytes_total) {
    if (!h) return set_err(EXB_ERR_ARG, "exb_reader_progress: null reader");
    // (streamed gzip / zstd / bzip2 / xz input is counted in text bytes against the compressed size: capped at the total)
    const int64_t total = h->r->bytes_total.load(), done = h->r->bytes_done.load();
    if (bytes_done) *bytes_done = total > 0 && done > total ? total : done;
    if (bytes_total) *bytes_total = total;
    return 0;
}

int exb_stats_snapshot(exb_scan_stats* out, int max, int* n) {
    if (!n || max < 0 || (max && !out)) return set_err(EXB_ERR_ARG, "exb_stats_snapshot: bad arguments");
    StatsRing& r = StatsRing::get();
    std::lock_guard<std::mutex> lk(r.mu);
    int k = 0;
    for (const exb_scan_stats& st : r.q) {
        if (k >= max) break;
        out[k++] = st;
    }
    *n = k;
    return 0;
}
void exb_stats_reset(void) {
    StatsRing& r = StatsRing::get();
    std::lock_guard<std::mutex> lk(r.mu);
    r.q.clear();
}

int exb_file_cache_state(const char* path) {
    struct stat sb;
    if (!path || stat(path, &sb) != 0 || !S_ISREG(sb.st_mode)) return set_err(EXB_ERR_IO, "exb_file_cache_state: cannot stat the file");
    std::shared_ptr<FileMap> fm = MapCache::get().find(sb);
    if (!fm) return 0;
    const int st = fm->reg.load();
    return st == 3 ? 0 : st;
}
int exb_reader_io_path(const exb_reader* h) { return h && h->r ? h->r->io_direct.load() : 0; }

void exb_reader_close(exb_reader* h) {
    if (!h) return;
    delete h->r;
    delete h;
}

}  // extern "C"
