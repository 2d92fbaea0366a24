// combine_records.cuh -- the per-shard records of a sharded COUNT (fastq_fused_candidates_kernel, fastq_scan.cu) -> the
// aggregates of the whole file.  Shared by the local combine kernel and the NVLink peer-memory exchange kernel
// (peer_exchange.cu), which runs it on every rank after ONE all-gather of the 256-byte records.
#pragma once
#include "exon_b200_internal.h"
#include "x87div.h"

namespace exb {

// Every shard's record -> the aggregates of the whole file (one thread; at most 16 shards).  ranges = {lo, hi, begin}
// per shard as in fastq_compose_prev_kernel.  total[0] passing records, [3] their Phred sums, [4] their lengths,
// [6] lines of the file mod 4, [7] 1 if a line start contradicts its phase.
__device__ __forceinline__ void fastq_combine_records(const long long* __restrict__ recs, const int64_t* __restrict__ ranges, int world,
                                      const exb_predicate* preds, int n_preds, long long* __restrict__ total) {
    long long cnt = 0, qs = 0, ql = 0;
    bool bad = false;
    unsigned long long lines = 0;
    long long ostart = ranges[0], os = 0;  // the line open at the first byte of shard j: start (file offset), byte sum so far
    // '@' / '+' flags of the byte at file offset p: the shard that holds the byte knows it -- as its own first byte, or as the
    // first byte of the line that is open after its last newline (p is always one of the two)
    auto flags_at = [&](long long p) -> uint32_t {
        for (int m = 0; m < world; m++)
            if (ranges[3 * m] <= p && p < ranges[3 * m + 1]) return (uint32_t)(p == ranges[3 * m] ? recs[32 * m + 8] : recs[32 * m + 3]) & 3u;
        return 0u;
    };
    for (int j = 0; j < world; j++) {
        const long long* r = recs + 32 * j;
        const long long lo = ranges[3 * j], begin = ranges[3 * j + 2];
        const int c = (int)(lines & 3);
        cnt += r[9 + 3 * c];
        qs += r[10 + 3 * c];
        ql += r[11 + 3 * c];
        bad = bad || (((unsigned long long)r[21] >> c) & 1ull) != 0;
        if (r[0] > 0) {  // the shard's first newline ends the open line
            const long long nl = lo + (r[5] - begin);
            long long len = nl - ostart;
            const long long cr = len > 0 ? r[6] : 0;
            len -= cr;
            if ((c & 1) == 0) {
                if (!(flags_at(ostart) & (c == 0 ? 2u : 1u))) bad = true;
            } else if (c == 3) {
                const long long q1 = r[7] + os - 13 * cr - 33 * len;
                bool ok = true;
                for (int i = 0; i < n_preds; i++) {
                    const exb_predicate p = preds[i];
                    ok = ok && (p.field == EXB_P_MEAN_QUALITY ? exb_mean_cmp(q1, (uint32_t)len, p.op, p.value) : exb_cmp((double)len, p.op, p.value));
                }
                if (ok) {
                    cnt += 1;
                    qs += q1;
                    ql += len;
                }
            }
            lines += (unsigned long long)r[0];
            ostart = lo + (r[1] - begin);
            os = r[2];
        } else {
            os += r[2];  // no newline: the whole shard belongs to the open line
        }
    }
    total[0] = cnt;
    total[1] = 0;
    total[2] = 0;
    total[3] = qs;
    total[4] = ql;
    total[5] = 0;
    total[6] = (long long)(lines & 3);
    total[7] = bad ? 1 : 0;
}

}  // namespace exb
