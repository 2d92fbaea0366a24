// x87div.h -- bit-exact emulation of `(double)((long double)sum / (long double)n)`
// as evaluated by x86-64 gcc (x87 80-bit extended, round-to-nearest-even twice).
//
// That expression is how DuckDB v0.8.1 finalises list_avg / avg over INTEGER
// (IntegerAverageOperationHugeint::Finalize), i.e. the reference's "mean
// quality" = list_avg(quality_score_string_to_list(quality_scores)).  The GPU
// has no 80-bit type, so the two roundings are done in integer arithmetic.
// Compiles for host (tests compare it with native long double) and device.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdbool.h>

#if defined(__CUDACC__)
#define EXB_HD __host__ __device__ __forceinline__
#define EXB_HD_COLD static __host__ __device__ __noinline__
#else
#define EXB_HD static inline
#define EXB_HD_COLD static
#endif

EXB_HD int exb_clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// 128-bit value hi:lo divided by a 32-bit divisor; quotient must fit 64 bits + 1.
EXB_HD void exb_div128_32(uint64_t hi, uint64_t lo, uint32_t n, uint64_t* q_hi, uint64_t* q_lo, uint32_t* rem) {
    uint32_t w[4] = {(uint32_t)(hi >> 32), (uint32_t)hi, (uint32_t)(lo >> 32), (uint32_t)lo};
    uint32_t qd[4];
    uint64_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t cur = (r << 32) | w[i];
        qd[i] = (uint32_t)(cur / n);
        r = cur % n;
    }
    *q_hi = ((uint64_t)qd[0] << 32) | qd[1];
    *q_lo = ((uint64_t)qd[2] << 32) | qd[3];
    *rem = (uint32_t)r;
}

EXB_HD double exb_x87_div(int64_t sum, uint32_t n) {
    if (sum == 0) return 0.0;
    const bool neg = sum < 0;
    uint64_t a = neg ? (uint64_t)(-(sum + 1)) + 1u : (uint64_t)sum;
    int la = 63 - exb_clz64(a), ln = 63 - exb_clz64((uint64_t)n);
    int e = la - ln;  // 2^(e-1) < a/n < 2^(e+1)
    uint64_t q64 = 0;
    uint32_t rem = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        int sh = 63 - e;  // 0 .. 94
        uint64_t hi, lo;
        if (sh == 0) {
            hi = 0;
            lo = a;
        } else if (sh < 64) {
            hi = a >> (64 - sh);
            lo = a << sh;
        } else {
            hi = a << (sh - 64);
            lo = 0;
        }
        uint64_t qh, ql;
        exb_div128_32(hi, lo, n, &qh, &ql, &rem);
        q64 = ql;  // qh == 0 by construction
        if (q64 >> 63) break;
        e -= 1;  // a/n was below 2^e: one more bit
    }
    // first rounding: 64-bit significand, ties to even
    uint64_t two_r = (uint64_t)rem * 2u;
    if (two_r > n || (two_r == n && (q64 & 1u))) {
        q64++;
        if (q64 == 0) {
            q64 = 1ull << 63;
            e += 1;
        }
    }
    // second rounding: 53-bit significand, ties to even
    uint64_t low = q64 & 0x7FFu, q53 = q64 >> 11;
    if (low > 0x400u || (low == 0x400u && (q53 & 1u))) q53++;
    double r = ldexp((double)q53, e - 52);
    return neg ? -r : r;
}

EXB_HD bool exb_cmp(double v, int op, double c) {
    switch (op) {
    case 0: return v > c;
    case 1: return v >= c;
    case 2: return v < c;
    case 3: return v <= c;
    case 4: return v == c;
    default: return v != c;
    }
}

// The rare part of exb_mean_cmp: sum/n is within a few ulp of c, so the two roundings of the x87 path decide.
// Kept out of line: it is ~200 instructions that almost no record executes.
EXB_HD_COLD bool exb_mean_cmp_close(int64_t sum, uint32_t n, int op, double c) {
    double q = (double)sum / (double)n;  // correctly rounded once; x87 may differ by 1 ulp
    double d = fabs(q - c);
    if (d > fabs(q) * 8.8817841970012523e-16) return exb_cmp(q, op, c);  // > 4 ulp away: same verdict
    return exb_cmp(exb_x87_div(sum, n), op, c);
}

// list_avg(...) <op> c for a list of n ints summing to `sum`; empty list = NULL = false.
EXB_HD bool exb_mean_cmp(int64_t sum, uint32_t n, int op, double c) {
    if (n == 0) return false;
    // Division-free verdict when sum/n is clearly on one side of c: D = sum - c*n has the sign of
    // sum/n - c, and |D| > 1e-14 |c n| puts the exact quotient more than 40 ulp away from c, which
    // neither of the two roundings can cross.  (t and d carry <= 2 ulp of error themselves.)
    const double t = c * (double)n, d = (double)sum - t;
    if (fabs(d) > fabs(t) * 1e-14) {
        switch (op) {
        case 0: case 1: return d > 0;
        case 2: case 3: return d < 0;
        case 4: return false;
        default: return true;
        }
    }
    return exb_mean_cmp_close(sum, n, op, c);
}
