// gen_common.h -- deterministic synthetic FASTA/FASTQ text,
// counter-based so any record can be produced on its own (host or device, any
// shard) and the bytes are identical everywhere.  Integer arithmetic only.
#pragma once
#include <stdint.h>

#include "exb_synth.h"

#if defined(__CUDACC__)
#define EXB_GHD __host__ __device__ inline
#else
#define EXB_GHD static inline
#endif

EXB_GHD uint64_t exb_mix(uint64_t z) {  // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
EXB_GHD uint64_t exb_hash(uint64_t seed, uint64_t rec, uint64_t stream, uint64_t idx) {
    return exb_mix(exb_mix(exb_mix(seed ^ (rec * 0xD6E8FEB86659FD93ull)) + stream * 0xA24BAED4963EE407ull) + idx);
}
// sum of four uniform bytes minus its mean: ~N(0, 147.8^2), range [-510, 510]
EXB_GHD int exb_zsum(uint32_t h) { return (int)(h & 0xFF) + (int)((h >> 8) & 0xFF) + (int)((h >> 16) & 0xFF) + (int)(h >> 24) - 510; }

EXB_GHD int exb_put_uint(uint8_t* out, uint64_t v) {  // decimal, returns digits (out may be null = count only)
    char tmp[20];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    if (out)
        for (int i = 0; i < n; i++) out[i] = (uint8_t)tmp[n - 1 - i];
    return n;
}
EXB_GHD int exb_put_str(uint8_t* out, const char* s) {
    int n = 0;
    while (s[n]) {
        if (out) out[n] = (uint8_t)s[n];
        n++;
    }
    return n;
}
EXB_GHD int exb_put_hex(uint8_t* out, uint64_t v, int digits) {
    for (int i = 0; i < digits; i++) {
        int d = (int)((v >> (4 * (digits - 1 - i))) & 0xF);
        if (out) out[i] = (uint8_t)(d < 10 ? '0' + d : 'a' + d - 10);
    }
    return digits;
}

EXB_GHD int64_t exb_rec_len(const exb_gen_params* p, uint64_t rec) {  // read / contig length
    if (p->len_max <= p->len_min) return p->len_min;
    uint64_t h = exb_hash(p->seed, rec, 1, 0);
    return p->len_min + (int64_t)(h % (uint64_t)(p->len_max - p->len_min + 1));
}

// Writes record `rec` (global index) at out (or only measures when out == null).  Returns its size.
EXB_GHD int64_t exb_gen_record(const exb_gen_params* p, uint64_t rec, uint8_t* out) {
    const int eol = p->crlf ? 2 : 1;
    int64_t o = 0;
#define EXB_EOL()                                 \
    do {                                          \
        if (out) {                                \
            if (p->crlf) out[o] = '\r';           \
            out[o + eol - 1] = '\n';              \
        }                                         \
        o += eol;                                 \
    } while (0)
    const int64_t L = exb_rec_len(p, rec);
    if (p->kind == EXB_GEN_FASTA) {
        o += exb_put_str(out ? out + o : 0, ">contig");
        o += exb_put_uint(out ? out + o : 0, rec);
        o += exb_put_str(out ? out + o : 0, " len=");
        o += exb_put_uint(out ? out + o : 0, (uint64_t)L);
        EXB_EOL();
        const int wrap = p->wrap > 0 ? p->wrap : 60;
        if (!out) {
            o += L + ((L + wrap - 1) / wrap) * eol;
            return o;
        }
        // per-contig GC fraction in [0.35, 0.65) as a 16-bit threshold; 1/1024 of bases are 'N'
        const uint32_t gcf = 22938u + (uint32_t)(exb_hash(p->seed, rec, 2, 0) % 19661u);
        for (int64_t i = 0; i < L; i++) {
            if ((i & 3) == 0) { /* one hash per 4 bases: 16 bits each */ }
            uint64_t h = exb_hash(p->seed, rec, 3, (uint64_t)(i >> 2));
            uint32_t u = (uint32_t)(h >> (16 * (i & 3))) & 0xFFFFu;
            uint8_t c;
            if ((u & 0x3FFu) == 0x3FFu && ((exb_hash(p->seed, rec, 4, (uint64_t)i) & 1u) != 0)) c = 'N';
            else {
                uint32_t u2 = (uint32_t)exb_mix(h + (uint64_t)i) & 1u;
                c = (u < gcf) ? (u2 ? 'G' : 'C') : (u2 ? 'A' : 'T');
            }
            out[o++] = c;
            if ((i + 1) % wrap == 0 || i + 1 == L) EXB_EOL();
        }
        return o;
    }
    // ---- FASTQ
    const bool ont = p->kind == EXB_GEN_ONT;
    const uint64_t hh = exb_hash(p->seed, rec, 5, 0);
    if (out) out[o] = '@';
    o += 1;
    if (!ont) {
        o += exb_put_str(out ? out + o : 0, "SIM:1:FC1:");
        o += exb_put_uint(out ? out + o : 0, 1 + (hh & 7));
        o += exb_put_str(out ? out + o : 0, ":");
        o += exb_put_uint(out ? out + o : 0, 1101 + ((hh >> 3) % 1578));
        o += exb_put_str(out ? out + o : 0, ":");
        o += exb_put_uint(out ? out + o : 0, 1000 + ((hh >> 16) % 29000));
        o += exb_put_str(out ? out + o : 0, ":");
        o += exb_put_uint(out ? out + o : 0, 1000 + ((hh >> 32) % 99000));
        o += exb_put_str(out ? out + o : 0, " ");
        o += exb_put_uint(out ? out + o : 0, 1 + ((hh >> 60) & 1));
        o += exb_put_str(out ? out + o : 0, ":N:0:ACGTACGT");
    } else {
        o += exb_put_hex(out ? out + o : 0, hh, 16);
        o += exb_put_hex(out ? out + o : 0, exb_mix(hh), 16);
        o += exb_put_str(out ? out + o : 0, " runid=");
        o += exb_put_hex(out ? out + o : 0, exb_mix(p->seed), 8);
        o += exb_put_str(out ? out + o : 0, " ch=");
        o += exb_put_uint(out ? out + o : 0, 1 + (exb_mix(hh + 1) % 512));
    }
    EXB_EOL();
    if (!out) {
        o += 2 * (L + eol) + 1 + eol;
        return o;
    }
    for (int64_t i = 0; i < L; i++) {
        uint64_t h = exb_hash(p->seed, rec, 6, (uint64_t)(i >> 5));
        out[o++] = (uint8_t)"ACGT"[(h >> (2 * (i & 31))) & 3];
    }
    EXB_EOL();
    out[o++] = '+';
    EXB_EOL();
    // fixed point (x1024): Illumina mean ~N(31,4), per base sd 5, clip [2,41]; ONT mean 12, sd 4, clip [1,40]
    const int zr = exb_zsum((uint32_t)(hh >> 20));
    const int mu = ont ? 12 * 1024 : 31 * 1024 + (4 * 1024 * zr) / 148;
    const int sd = ont ? 4 : 5, qlo = ont ? 1 : 2, qhi = ont ? 40 : 41;
    for (int64_t i = 0; i < L; i++) {
        uint64_t h = exb_hash(p->seed, rec, 7, (uint64_t)(i >> 1));
        int z = exb_zsum((uint32_t)(h >> (32 * (i & 1))));
        int v = mu + (sd * 1024 * z) / 148 + 512 + 65536;
        int q = (v >> 10) - 64;
        q = q < qlo ? qlo : (q > qhi ? qhi : q);
        out[o++] = (uint8_t)(33 + q);
    }
    EXB_EOL();
    return o;
#undef EXB_EOL
}
