// TEST INFRASTRUCTURE ONLY -- the DEFLATE / CRC-32 core of the device inflater (exon_duckdb_b200/csrc/inflate_core.cuh),
// compiled for the host with ONE lane, so that the algorithm can be compared with zlib (the oracle for this row: zlib
// through Python's `zlib` module) in the CPU test tier.  Nothing in the product path links or loads this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../exon_duckdb_b200/csrc/inflate_core.cuh"

extern "C" {

// inflate one raw DEFLATE stream; returns ifl::Err
int ifl_host_inflate(const uint8_t* in, int clen, uint8_t* out, int isize, int misalign) {
    // the device reader loads whole words up to 12 bytes past the payload and from the aligned word in front of it
    uint8_t* buf = (uint8_t*)calloc((size_t)clen + 64, 1);
    uint8_t* p = buf + 16 + (misalign & 3);
    memcpy(p, in, (size_t)clen);
    ifl::Tables* T = new ifl::Tables();
    const int rc = ifl::inflate_block<1>(p, clen, out, isize, *T, 0);
    delete T;
    free(buf);
    return rc;
}

// CRC-32 the way the kernel computes it: 32 slices, combined in order
uint32_t ifl_host_crc32(const uint8_t* p, int n) {
    uint32_t tab[256];
    for (uint32_t i = 0; i < 256; i++) tab[i] = ifl::crc_table_entry(i);
    const int per = (n + 31) >> 5;
    uint32_t total = 0;
    const uint32_t x_per = ifl::crc_shift_bytes((uint32_t)per);
    for (int lane = 0; lane < 32; lane++) {
        const int lo = lane * per < n ? lane * per : n;
        const int mine = n - lo < per ? n - lo : per;
        const uint32_t c = ifl::crc_bytes(p + lo, mine, tab);
        if (lane == 0) {
            total = c;
            continue;
        }
        if (mine == 0) break;
        total = ifl::crc_multmodp(mine == per ? x_per : ifl::crc_shift_bytes((uint32_t)mine), total) ^ c;
    }
    return total;
}
}
