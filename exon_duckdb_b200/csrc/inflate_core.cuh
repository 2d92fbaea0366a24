// DEFLATE (RFC 1951) decoder for ONE compressed block of a BGZF file (SAM spec 4.1: a gzip member of <= 64 KiB whose extra
// field carries its own size), written so that one WARP decodes one block:
//   * the bit reader, the Huffman decode loop and the dynamic-header parse run uniformly (every lane holds the same state;
//     an instruction costs one issue slot whether 1 or 32 lanes take part), lane 0 stores the literals;
//   * the table fill, the LZ77 match copy and stored blocks are spread over the lanes;
//   * the block's own output in global memory is the LZ77 window (a BGZF block never refers to an earlier block).
// The same source compiles for the host with ONE lane (IFL_SYNC a no-op): oracle/inflate_host.cpp wraps that build so the
// algorithm can be checked against zlib without a GPU (tests/test_inflate.py).  That host build is test infrastructure;
// the product path is the kernel in inflate.cu.
//
// Replaces, for BGZF input: the streaming GzipDecoder the reference wraps around the file (rust/src/arrow_reader.rs:60-91,
// datafusion FileCompressionType::GZIP) -- SURVEY 8(f) rank 1 "BGZF ... GPU inflate later".  The reference's decoder is a
// third-party dependency that is not in its tree (rust/Cargo.lock: async-compression 0.4.0 -> flate2 1.0.26 ->
// miniz_oxide 0.7.1); the algorithm is the published one, RFC 1951 (DEFLATE) inside RFC 1952 (gzip) members, and parity
// is anchored on zlib -- the implementation every gzip file is defined against -- and on the reference's own .gz fixtures.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define IFL_HD __host__ __device__ __forceinline__
#else
#define IFL_HD inline
#endif
#ifdef __CUDA_ARCH__
#define IFL_SYNC() __syncwarp()
#define IFL_LD32(p) __ldg(p)
#else
#define IFL_SYNC() ((void)0)
#define IFL_LD32(p) (*(p))
#endif

namespace ifl {

#ifndef IFL_LB  // (the host test build also compiles a variant with tiny tables: every code word takes the bit-by-bit path)
#define IFL_LB 9   // 9 + 8 bits: 4 KB of tables per warp, 6 CTAs of 8 warps per SM.  Measured against 10 + 8 bits at 4 CTAs
#define IFL_DB 8   // (31.5 ms per 1.12 GB of text) and 9 + 7 bits at 8 CTAs (28.7 ms): 28.2 ms (profiles/round2_inflate_ab.txt)
#endif
constexpr int LB = IFL_LB;  // index bits of the literal / length table
constexpr int DB = IFL_DB;  // index bits of the distance table
constexpr int CLB = 7;   // the code-length code: at most 7 bits, always resolved by its table
enum Kind { K_SLOW = 0, K_LIT = 1, K_BASE = 2, K_EOB = 3 };
enum Mode { M_LITLEN = 0, M_DIST = 1, M_CL = 2 };
enum Err {
    OK = 0,
    E_BTYPE = 1,       // reserved block type 3
    E_STORED = 2,      // LEN / NLEN mismatch
    E_HEADER = 3,      // dynamic header: bad counts or repeat without a previous length
    E_OVERSUB = 4,     // over-subscribed code
    E_CODE = 5,        // a bit pattern that is no code word / an invalid symbol
    E_DIST = 6,        // distance reaches in front of the block
    E_OUTPUT = 7,      // more output than the block's ISIZE says / less
    E_INPUT = 8,       // the stream runs past the compressed payload
    E_CRC = 9,         // CRC-32 of the output differs from the trailer
};

// entry: bits 0-4 code length, 5-7 kind, 8-11 extra bits, 16-30 literal byte | base length | base distance | symbol (M_CL);
// bit 31 = literal of the literal / length code (the decode loop's first test is the sign)
struct Tables {
    uint32_t lit[1 << LB];
    uint32_t dist[1 << DB];     // (the code-length code's table lives in its first 1 << CLB entries while a header is read)
    uint16_t lit_sorted[288];   // symbols in canonical order (length, then value): code words longer than the table's index
    uint16_t dist_sorted[32];
    uint16_t lit_count[16];     // number of code words per length
    uint16_t dist_count[16];
    uint8_t lens[320];          // code lengths of the header being read (literal / length codes, then distance codes)
};

IFL_HD uint32_t make_entry(int mode, int sym, int len) {
    if (mode == M_CL) return (uint32_t)len | (K_LIT << 5) | ((uint32_t)sym << 16);
    if (mode == M_DIST) {
        if (sym >= 30) return 0u;
        const int xb = sym < 4 ? 0 : (sym >> 1) - 1;
        const uint32_t base = sym < 4 ? 1u + sym : 1u + ((2u + (sym & 1)) << xb);
        return (uint32_t)len | (K_BASE << 5) | ((uint32_t)xb << 8) | (base << 16);
    }
    if (sym < 256) return 0x80000000u | (uint32_t)len | (K_LIT << 5) | ((uint32_t)sym << 16);
    if (sym == 256) return (uint32_t)len | (K_EOB << 5);
    if (sym >= 286) return 0u;
    const int idx = sym - 257;
    int xb = 0;
    uint32_t base = 258u;
    if (sym != 285) {
        xb = idx < 8 ? 0 : (idx >> 2) - 1;
        base = idx < 8 ? 3u + idx : 3u + ((4u + (idx & 3)) << xb);
    }
    return (uint32_t)len | (K_BASE << 5) | ((uint32_t)xb << 8) | (base << 16);
}

IFL_HD uint32_t bit_reverse(uint32_t v, int n) {  // the low n bits of v, reversed
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// Canonical Huffman code from n code lengths: the index table (code words of <= TB bits, every slot they are a prefix of),
// the symbols in canonical order and the per-length counts (longer code words are decoded from those, bit by bit).
// Incomplete codes are accepted (their unused bit patterns are errors only if the stream contains one), over-subscribed
// ones are not.  Uniform part by lane 0, table fill over the lanes.
template <int TB>
IFL_HD int build_table(const uint8_t* lens, int n, uint32_t* table, uint16_t* sorted, uint16_t* count, int mode, int lane, int lanes) {
    uint16_t offs[16];
    uint16_t next_code[16];
    IFL_SYNC();  // (the table's earlier readers are done)
    if (lane == 0) {
        for (int l = 0; l < 16; l++) count[l] = 0;
        for (int s = 0; s < n; s++) count[lens[s]]++;
    }
    for (int j = lane; j < (1 << TB); j += lanes) table[j] = 0u;
    IFL_SYNC();
    int left = 1, code = 0;
    offs[0] = 0;
    offs[1] = 0;
    next_code[0] = 0;
    for (int l = 1; l < 16; l++) {
        left = (left << 1) - (int)count[l];
        if (left < 0) return E_OVERSUB;
        code = (code + (l > 1 ? (int)count[l - 1] : 0)) << 1;
        next_code[l] = (uint16_t)code;
        if (l < 15) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    }
    const int total = offs[15] + count[15];
    if (lane == 0) {
        uint16_t o[16];
        for (int l = 0; l < 16; l++) o[l] = offs[l];
        for (int s = 0; s < n; s++)
            if (lens[s]) sorted[o[lens[s]]++] = (uint16_t)s;
    }
    IFL_SYNC();
    for (int k = lane; k < total; k += lanes) {
        const int s = sorted[k];
        const int len = lens[s];
        if (len > TB) continue;
        const uint32_t cw = (uint32_t)next_code[len] + (uint32_t)(k - offs[len]);
        const uint32_t e = make_entry(mode, s, len);
        for (uint32_t j = bit_reverse(cw, len); j < (1u << TB); j += 1u << len) table[j] = e;
    }
    IFL_SYNC();
    return OK;
}

// a code word that the index table does not hold: canonical decode, one bit at a time, from the peeked bits
IFL_HD uint32_t decode_slow(uint32_t v, const uint16_t* sorted, const uint16_t* count, int mode) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)((v >> (len - 1)) & 1u);
        const int c = count[len];
        if (code - c < first) return make_entry(mode, sorted[index + (code - first)], len);
        index += c;
        first = (first + c) << 1;
        code <<= 1;
    }
    return 0u;
}

struct BitReader {
    const uint32_t* wp;     // next word to load
    const uint32_t* limit;  // first word that must not be loaded (a corrupt stream may ask for any number of bits: it gets zeros)
    uint32_t w0, w1, w2;
    int bo;                 // bits of w0 already consumed
    IFL_HD void init(const uint8_t* p, const uint8_t* readable_end) {
        limit = (const uint32_t*)((uintptr_t)readable_end & ~(uintptr_t)3);
        const uintptr_t a = (uintptr_t)p;
        wp = (const uint32_t*)(a & ~(uintptr_t)3);
        bo = (int)(a & 3) * 8;
        w0 = IFL_LD32(wp);
        w1 = IFL_LD32(wp + 1);
        w2 = IFL_LD32(wp + 2);
        wp += 3;
    }
    IFL_HD uint32_t peek() const {  // the next 32 bits
#ifdef __CUDA_ARCH__
        return __funnelshift_r(w0, w1, bo);
#else
        return bo ? (w0 >> bo) | (w1 << (32 - bo)) : w0;
#endif
    }
    IFL_HD void skip(int n) {  // n <= 32
        bo += n;
        if (bo >= 32) {
            bo -= 32;
            w0 = w1;
            w1 = w2;
            w2 = wp < limit ? IFL_LD32(wp) : 0u;  // needed only after the NEXT crossing: its latency hides behind the symbols in between
            wp++;
        }
    }
    IFL_HD const uint8_t* byte_ptr() const { return (const uint8_t*)(wp - 3) + ((bo + 7) >> 3); }  // first byte not touched yet
};

// Inflate one raw DEFLATE stream of `clen` bytes at `in` into exactly `isize` bytes at `out`.  The caller guarantees 16
// readable bytes behind in + clen (the reader loads whole words ahead).  Returns an Err.
template <int LANES>
IFL_HD int inflate_block(const uint8_t* in, int clen, uint8_t* out, int isize, Tables& T, int lane) {
    constexpr int lanes = LANES;
    BitReader br;
    br.init(in, in + clen + 16);
    int pos = 0;
    for (;;) {
        if (br.byte_ptr() > in + clen) return E_INPUT;
        uint32_t v = br.peek();
        const int bfinal = (int)(v & 1u), btype = (int)((v >> 1) & 3u);
        br.skip(3);
        if (btype == 3) return E_BTYPE;
        if (btype == 0) {
            br.skip((8 - (br.bo & 7)) & 7);
            v = br.peek();
            const uint32_t len = v & 0xFFFFu, nlen = v >> 16;
            br.skip(32);
            if ((len ^ nlen) != 0xFFFFu) return E_STORED;
            const uint8_t* src = br.byte_ptr();
            if (src + len > in + clen) return E_INPUT;
            if (pos + (int)len > isize) return E_OUTPUT;
            for (int i = lane; i < (int)len; i += lanes) out[pos + i] = src[i];
            pos += (int)len;
            br.init(src + len, in + clen + 16);
        } else {
            int err = OK;
            if (btype == 1) {
                if (lane == 0) {
                    for (int s = 0; s < 144; s++) T.lens[s] = 8;
                    for (int s = 144; s < 256; s++) T.lens[s] = 9;
                    for (int s = 256; s < 280; s++) T.lens[s] = 7;
                    for (int s = 280; s < 288; s++) T.lens[s] = 8;
                    for (int s = 0; s < 32; s++) T.lens[288 + s] = 5;
                }
                IFL_SYNC();
                err = build_table<LB>(T.lens, 288, T.lit, T.lit_sorted, T.lit_count, M_LITLEN, lane, lanes);
                if (err == OK) err = build_table<DB>(T.lens + 288, 32, T.dist, T.dist_sorted, T.dist_count, M_DIST, lane, lanes);
                if (err != OK) return err;
            } else {
                v = br.peek();
                const int hlit = (int)(v & 31u) + 257, hdist = (int)((v >> 5) & 31u) + 1, hclen = (int)((v >> 10) & 15u) + 4;
                br.skip(14);
                if (hlit > 286 || hdist > 30) return E_HEADER;
                const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                IFL_SYNC();
                if (lane == 0)
                    for (int i = 0; i < 19; i++) T.lens[i] = 0;
                IFL_SYNC();
                for (int i = 0; i < hclen; i++) {
                    const uint32_t l = br.peek() & 7u;
                    br.skip(3);
                    if (lane == 0) T.lens[order[i]] = (uint8_t)l;
                }
                IFL_SYNC();
                // (the code-length code borrows the distance table and the distance code's canonical arrays)
                err = build_table<CLB>(T.lens, 19, T.dist, T.dist_sorted, T.dist_count, M_CL, lane, lanes);
                if (err != OK) return err;
                int i = 0, prev = -1;
                const int want = hlit + hdist;
                while (i < want) {
                    v = br.peek();
                    const uint32_t e = T.dist[v & ((1u << CLB) - 1u)];
                    if (e == 0u) return E_CODE;
                    const int nb = (int)(e & 31u), sym = (int)(e >> 16);
                    int rep = 1, val = sym;
                    if (sym < 16) {
                        br.skip(nb);
                        prev = sym;
                    } else if (sym == 16) {
                        if (prev < 0) return E_HEADER;
                        rep = 3 + (int)((v >> nb) & 3u);
                        val = prev;
                        br.skip(nb + 2);
                    } else if (sym == 17) {
                        rep = 3 + (int)((v >> nb) & 7u);
                        val = 0;
                        prev = 0;
                        br.skip(nb + 3);
                    } else {
                        rep = 11 + (int)((v >> nb) & 127u);
                        val = 0;
                        prev = 0;
                        br.skip(nb + 7);
                    }
                    if (i + rep > want) return E_HEADER;
                    if (lane == 0)
                        for (int k = 0; k < rep; k++) T.lens[i + k] = (uint8_t)val;
                    i += rep;
                }
                IFL_SYNC();
                if (T.lens[256] == 0) return E_HEADER;  // no end-of-block code
                err = build_table<LB>(T.lens, hlit, T.lit, T.lit_sorted, T.lit_count, M_LITLEN, lane, lanes);
                if (err == OK) err = build_table<DB>(T.lens + hlit, hdist, T.dist, T.dist_sorted, T.dist_count, M_DIST, lane, lanes);
                if (err != OK) return err;
            }
            // ---- the symbols of this block.  The kernel is bound by instruction issue (profiles/round2_inflate_v1_ncu.txt:
            // 72 % issue utilisation, 27 warp instructions per byte of text), so the loop is written for few instructions:
            // a literal is recognised by the entry's sign, stored by EVERY lane (same byte, same address: one transaction
            // and no divergent branch; each lane then reads its own store back without a barrier), and bounds are tested
            // once per symbol.
            const uint32_t* __restrict__ lit = T.lit;
            const uint32_t* __restrict__ dtab = T.dist;
            for (;;) {
                v = br.peek();
                uint32_t e = lit[v & ((1u << LB) - 1u)];
                if ((int32_t)e < 0) {  // literal
                    if (pos >= isize) return E_OUTPUT;
                    // ... and, more often than not, another one: the peeked word still holds its code (two table-resolved
                    // code words are <= 2 x LB bits), so it costs one more look-up instead of one more trip round the loop
                    const int nb1 = (int)(e & 31u);
                    const uint32_t e2 = lit[(v >> nb1) & ((1u << LB) - 1u)];
                    out[pos] = (uint8_t)(e >> 16);
                    if ((int32_t)e2 < 0 && pos + 1 < isize) {
                        out[pos + 1] = (uint8_t)(e2 >> 16);
                        pos += 2;
                        br.skip(nb1 + (int)(e2 & 31u));
                        continue;
                    }
                    pos++;
                    br.skip(nb1);
                    continue;
                }
                if ((e & 0xE0u) == 0u) {  // a code word longer than the table's index
                    e = decode_slow(v, T.lit_sorted, T.lit_count, M_LITLEN);
                    if (e == 0u) return E_CODE;
                    if ((int32_t)e < 0) {
                        if (pos >= isize) return E_OUTPUT;
                        out[pos++] = (uint8_t)(e >> 16);
                        br.skip((int)(e & 31u));
                        continue;
                    }
                }
                const int nb = (int)(e & 31u);
                if (((e >> 5) & 7u) == K_EOB) {
                    br.skip(nb);
                    break;
                }
                int xb = (int)((e >> 8) & 15u);
                const int len = (int)(e >> 16) + (int)((v >> nb) & ((1u << xb) - 1u));
                br.skip(nb + xb);
                v = br.peek();
                e = dtab[v & ((1u << DB) - 1u)];
                if ((e & 0xE0u) == 0u) {
                    e = decode_slow(v, T.dist_sorted, T.dist_count, M_DIST);
                    if (e == 0u) return E_CODE;
                }
                const int nd = (int)(e & 31u);
                xb = (int)((e >> 8) & 15u);
                const int dist = (int)(e >> 16) + (int)((v >> nd) & ((1u << xb) - 1u));
                br.skip(nd + xb);
                if (dist > pos) return E_DIST;
                if (pos + len > isize) return E_OUTPUT;
                IFL_SYNC();  // the bytes the match refers to were stored by other lanes
                const uint8_t* from = out + pos - dist;
                uint8_t* to = out + pos;
                if (dist >= len) {
                    if (LANES == 32 && len <= 32) {
                        if (lane < len) to[lane] = from[lane];
                    } else {
                        for (int i = lane; i < len; i += lanes) to[i] = from[i];
                    }
                } else if (dist == 1) {
                    const uint8_t b = from[0];
                    for (int i = lane; i < len; i += lanes) to[i] = b;
                } else {  // the match overlaps its own output: byte i repeats byte i mod dist of what was there before
                    for (int i = lane; i < len; i += lanes) to[i] = from[i % dist];
                }
                pos += len;
            }
        }
        if (bfinal) break;
    }
    if (pos != isize) return E_OUTPUT;
    if (br.byte_ptr() > in + clen) return E_INPUT;
    return OK;
}

// ---------------------------------------------------------------- CRC-32 (gzip trailer; RFC 1952 8.1.1, zlib's crc32_combine)
constexpr uint32_t CRC_POLY = 0xEDB88320u;
IFL_HD uint32_t crc_table_entry(uint32_t i) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
    return c;
}
// a(x) * b(x) mod p(x), reflected representation (bit 31 = x^0)
IFL_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod p(x): what a CRC is multiplied by when n bytes follow it
IFL_HD uint32_t crc_shift_bytes(uint32_t n) {
    uint32_t p = 0x80000000u;       // 1
    uint32_t sq = 0x00800000u;      // x^8 (bit 31 - 8)
    while (n) {
        if (n & 1u) p = crc_multmodp(sq, p);
        sq = crc_multmodp(sq, sq);
        n >>= 1;
    }
    return p;
}
// CRC-32 of n bytes: every lane takes a slice of `per` bytes (table look-ups), the slice CRCs are combined in order.
// `tab` = the 256-entry table (crc_table_entry).  On the device the combination runs over warp shuffles (inflate.cu);
// this is the slice part and the host's one-lane whole.
IFL_HD uint32_t crc_bytes(const uint8_t* p, int n, const uint32_t* tab) {
    uint32_t c = 0xFFFFFFFFu;
    for (int i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

}  // namespace ifl
