/*
 * exb_synth.h -- deterministic synthetic FASTA / FASTQ inputs.
 * TEST AND BENCH TOOLING, not part of the product library: libexb_synth.so is loaded by tests/, bench.py and
 * scripts/ only.  Counter-based RNG: any record can be produced on its own, on the host or on the device, and the
 * bytes are identical everywhere.
 */
#ifndef EXB_SYNTH_H
#define EXB_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#define EXB_API __attribute__((visibility("default")))
#else
#define EXB_API
#endif
#define EXB_GEN_ILLUMINA 2 /* C2/C5: 150 bp reads, '@SIM:1:FC1:lane:tile:x:y 1:N:0:ACGTACGT' */
#define EXB_GEN_ONT 4      /* C4: 10-50 kb reads                                               */
#define EXB_GEN_FASTA 1    /* C1/C3: wrapped FASTA                                              */
typedef struct exb_gen_params {
    int32_t kind;
    uint64_t seed;
    int64_t n_records;
    int64_t first_record; /* global index of record 0 of this call (sharded generation) */
    int32_t len_min, len_max; /* read / contig length range (inclusive) */
    int32_t wrap;         /* FASTA line width */
    int32_t crlf;         /* 1 = CRLF line ends */
} exb_gen_params;
/* Size in bytes of the text the parameters describe (host computation, exact). */
EXB_API int64_t exb_gen_size(const exb_gen_params *p);
/* The same size computed on the device (large record counts). Synchronous. */
EXB_API int exb_gen_size_device(const exb_gen_params *p, int64_t *size_out, void *stream);
/* Generate on the device: d_out must hold exb_gen_size(p) bytes (+16 slack). Synchronous. */
EXB_API int exb_gen_device(const exb_gen_params *p, void *d_out, int64_t cap, void *stream);
/* Generate on the host (same bytes), no GPU needed. */
EXB_API int exb_gen_host(const exb_gen_params *p, void *out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
