/* mk_synth.h — deterministic synthetic reads for the BASELINE configs.
 * Benchmark / test support, not part of the matching product. Counter-based: every byte of every
 * read is a pure function of (seed, read index, position), so the host and the device produce
 * identical data and any sub-range can be generated on its own.
 *
 *   base(r, i)   = 2 bits of splitmix64(seed + r * words_per_read + i / 32), iid uniform ACGT
 *   query j      = the k bases at position (j mod (read_len-k+1)) of read (j * 99991 mod n_reads),
 *                  taken from base(), i.e. before planting / N runs (so queries never contain N)
 *   read r       : with probability plant_per_65536/65536 one query (forward or reverse complement,
 *                  coin flip) is written at a random position; with probability
 *                  nrun_per_65536/65536 a run of 1..10 'N' is written afterwards.
 */
#ifndef MK_SYNTH_H
#define MK_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t seed;
    uint64_t n_reads;          /* size of the whole data set (defines the query sampling) */
    uint32_t read_len;
    uint32_t k;
    uint32_t n_queries;
    uint32_t plant_per_65536;  /* 655 ~ 1 % */
    uint32_t nrun_per_65536;   /* 328 ~ 0.5 % */
    uint32_t reserved;
} mks_params;

/* out: n_queries * k bytes */
int mks_queries(const mks_params* p, uint8_t* out);
/* Reads [r0, r1) into host memory. enc 0: ASCII, 1: BAM 4-bit (read_len must be even).
 * off (may be NULL): r1-r0+1 unit offsets relative to the first read. */
int mks_fill_host(const mks_params* p, const uint8_t* queries, uint64_t r0, uint64_t r1, int enc, uint8_t* out,
                  uint64_t* off);
/* Same into device memory, on `stream` (a cudaStream_t, may be NULL). d_queries: device copy of
 * the query table. */
int mks_fill_device(const mks_params* p, const uint8_t* d_queries, uint64_t r0, uint64_t r1, int enc, void* d_out,
                    uint64_t* d_off, void* stream);
const char* mks_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
