/* merkurio_cuda.h — C ABI of the B200 k-mer / multi-pattern matching engine.
 *
 * This is the drop-in boundary for MerKurio's matching hot path. The reference has no FFI: its
 * matchers are in-process Rust objects used from three inner loops. Each entry point below names
 * the reference interface it replaces (paths relative to the reference repository):
 *
 *   mk_engine_create   <- matcher construction: BNDMq::new per pattern (src/pattern_matching.rs:61-78,
 *                         masks from src/pattern_preprocessing.rs:24-43) and
 *                         AhoCorasick::builder().kind(DFA).ascii_case_insensitive(b).build(list)
 *                         (src/cmd_extract.rs:259-277, src/cmd_tag.rs:234-252)
 *   mk_scan_submit /   <- the per-record search calls: BNDMq::find_iter / find_match
 *   mk_scan_wait /        (src/pattern_matching.rs:128-140,165-209) and
 *   mk_scan_host /        AhoCorasick::find_overlapping_iter, as used by the extract loops
 *   mk_scan_device /      (src/cmd_extract.rs:321-406 single, :463-607 paired) and by
 *   mk_scan_device_submit process_record (src/cmd_tag.rs:387-443)
 *   mk_result          <- what those loops consume: found_occ (cmd_extract.rs:323,400),
 *                         kmers_found (cmd_tag.rs:387,398,429,439) and the (pattern, start) stream
 *                         handed to the loggers (src/logger.rs:41-60,108-133)
 *
 * The calls are batch-granular because a per-record call cannot feed a GPU. A batch is a set of
 * records whose sequence bytes are concatenated; record r occupies units [off[r], off[r]+len(r))
 * where len(r) = lens[r] if lens != NULL else off[r+1]-off[r]. A "unit" is one byte (= one base)
 * for MK_ENC_ASCII and one base (= one nibble, first base in the high nibble of a byte, record
 * starts on even unit offsets) for MK_ENC_BAM4.
 *
 * Results are exact: a hit (record, start, pattern) is reported iff the pattern's bytes equal the
 * record's bytes at [start, start+len) (ASCII letters compared case-folded when
 * mk_config.case_insensitive != 0) — the same set both reference matchers produce. Hits are sorted
 * by (record, start+len, start, pattern): the order of AhoCorasick::find_overlapping_iter. The
 * BNDMq order of the reference logs (pattern-major) is a host-side regrouping of that list.
 *
 * Conventions: every function returns 0 on success and a negative mk_status on failure;
 * mk_last_error() gives the message (thread-local). Nothing throws across the boundary. The engine
 * owns every buffer it hands out; views in mk_result stay valid until the next submit/scan on the
 * same slot (or engine, for mk_scan_device) or until mk_engine_destroy. One host thread drives one
 * engine; different engines (one per GPU) may be driven from different threads.
 * There is no CPU fallback: without a CUDA device mk_engine_create fails with MK_ERR_CUDA.
 */
#ifndef MERKURIO_CUDA_H
#define MERKURIO_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mk_engine mk_engine;

typedef enum {
    MK_OK = 0,
    MK_ERR_INVALID = -1,     /* bad argument */
    MK_ERR_EMPTY_PATTERN = -2, /* PatternError::EmptyPattern, src/pattern_matching.rs:32-33 */
    MK_ERR_NO_PATTERNS = -3, /* "No k-mers found in file or provided sequence.", src/helpers.rs:128-130 */
    MK_ERR_CUDA = -4,        /* CUDA runtime error / no device */
    MK_ERR_NOMEM = -5,
    MK_ERR_CAPACITY = -6,    /* batch larger than the slot was configured for */
    MK_ERR_STATE = -7        /* wait without submit, bad slot, ... */
} mk_status;

/* The query list after src/helpers.rs:76-133 (parse_pattern_list): sorted bytewise, unique, no
 * empty entries. Pattern index == index in this list. off has n+1 entries. */
typedef struct {
    const uint8_t* bytes;
    const uint32_t* off;
    uint32_t n;
} mk_patterns;

typedef struct {
    int32_t device;              /* CUDA device ordinal */
    int32_t case_insensitive;    /* -I: ascii_case_insensitive(true), src/cmd_extract.rs:263 */
    uint32_t n_slots;            /* pinned/device staging slots for submit/wait (0 = none; >=2 to double-buffer) */
    uint32_t max_batch_records;  /* per slot */
    uint64_t max_batch_bytes;    /* per slot, bytes of sequence storage (ASCII: bases; BAM4: bases/2) */
    uint64_t hit_capacity;       /* initial hit-list capacity per slot (grown on overflow; 0 = default) */
} mk_config;

typedef enum { MK_ENC_ASCII = 0, MK_ENC_BAM4 = 1 } mk_encoding;

typedef enum {
    MK_MODE_FLAG = 0,        /* 1 bit / record: any hit (extract without -l/-j) */
    MK_MODE_PATTERN_SET = 1, /* distinct (record, pattern) pairs, sorted (tag without -l/-j) */
    MK_MODE_ALL_HITS = 2     /* every (record, start, pattern), AC order (-l / -j) */
} mk_mode;

typedef struct {
    uint32_t record;  /* index of the record in the batch */
    uint32_t start;   /* zero-based start inside the record (0 in MK_MODE_PATTERN_SET) */
    uint32_t pattern; /* index into mk_patterns */
    uint32_t len;     /* pattern length */
} mk_hit;

typedef struct {
    const uint64_t* record_flags; /* host bitmap, bit (r & 63) of word r >> 6 set iff record r has >= 1 hit */
    uint32_t n_records;
    uint32_t reserved;
    const mk_hit* hits;           /* host array, sorted; NULL when n_hits == 0 or mode == FLAG */
    uint64_t n_hits;
    uint64_t bases_scanned;
    uint64_t device_ns;           /* CUDA-event time of the device work for this batch (kernels only) */
    uint64_t scan_ns;             /* CUDA-event time of the scan kernel alone */
    uint64_t verify_ns;           /* CUDA-event time of the candidate verification kernel */
    uint64_t n_candidates;        /* seeds that passed both filters and were verified */
    uint32_t n_rescans;           /* >0 if the hit list overflowed and the batch was scanned again */
    uint32_t reserved2;
    const uint64_t* d_record_flags; /* device copies of the above (valid like the host views) */
    const mk_hit* d_hits;
} mk_result;

/* What the table builder chose; for logs, tests and the roofline report. */
typedef struct {
    uint32_t n_patterns, min_len, max_len;
    uint32_t seed_q[2], seed_d[2];     /* per encoding (index = mk_encoding); 0 = tables not built yet */
    uint32_t n_seeds[2];               /* distinct seed codes */
    uint32_t filter_log2_bits[2];      /* first-level filter, L2-resident plain bitmap (stride 16): log2 of its bits (else 0) */
    uint32_t filter_hashes[2];         /* bits tested per probe */
    uint64_t filter_bytes[2];          /* size of the first-level filter */
    uint32_t filter_in_smem[2];        /* 1: bitmap staged in shared memory, 0: L2-resident */
    uint64_t table_bytes[2];           /* cuckoo seed table + postings + pattern bytes */
    uint32_t sm_count;
    uint32_t features;                 /* MK_FEATURE_* of the ASCII tables in bits 0..7, of the BAM4 tables in bits 8..15 */
} mk_engine_info;
#define MK_FEATURE_DUAL8 1u            /* stride-8 scan with the L2-resident dual-key filter (large query sets) */
#define MK_FEATURE_GATE 2u             /* the query alphabet allows an alphabet gate (windows holding a byte no pattern contains are not probed) */

/* Layout contract with bindings in other languages (the #[repr(C)] structs): sizes and offsets
 * on the LP64 targets this library is built for. A change here is an ABI break. */
#if defined(__cplusplus)
#define MK_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define MK_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
#include <stddef.h>
MK_STATIC_ASSERT(sizeof(mk_patterns) == 24 && offsetof(mk_patterns, off) == 8 && offsetof(mk_patterns, n) == 16, "mk_patterns layout");
MK_STATIC_ASSERT(sizeof(mk_config) == 32 && offsetof(mk_config, n_slots) == 8 && offsetof(mk_config, max_batch_records) == 12 &&
                 offsetof(mk_config, max_batch_bytes) == 16 && offsetof(mk_config, hit_capacity) == 24, "mk_config layout");
MK_STATIC_ASSERT(sizeof(mk_hit) == 16 && offsetof(mk_hit, start) == 4 && offsetof(mk_hit, pattern) == 8 && offsetof(mk_hit, len) == 12, "mk_hit layout");
MK_STATIC_ASSERT(sizeof(mk_result) == 96 && offsetof(mk_result, n_records) == 8 && offsetof(mk_result, hits) == 16 && offsetof(mk_result, n_hits) == 24 &&
                 offsetof(mk_result, bases_scanned) == 32 && offsetof(mk_result, device_ns) == 40 && offsetof(mk_result, scan_ns) == 48 &&
                 offsetof(mk_result, verify_ns) == 56 && offsetof(mk_result, n_candidates) == 64 && offsetof(mk_result, n_rescans) == 72 &&
                 offsetof(mk_result, d_record_flags) == 80 && offsetof(mk_result, d_hits) == 88, "mk_result layout");
MK_STATIC_ASSERT(sizeof(mk_engine_info) == 104 && offsetof(mk_engine_info, seed_q) == 12 && offsetof(mk_engine_info, filter_bytes) == 56 &&
                 offsetof(mk_engine_info, filter_in_smem) == 72 && offsetof(mk_engine_info, table_bytes) == 80 && offsetof(mk_engine_info, sm_count) == 96 &&
                 offsetof(mk_engine_info, features) == 100, "mk_engine_info layout");

int mk_engine_create(const mk_patterns* patterns, const mk_config* config, mk_engine** out);
void mk_engine_destroy(mk_engine* e);
int mk_engine_get_info(mk_engine* e, mk_engine_info* out);
/* Name and flavour of the scan kernel the tables of `enc` run on ("" until a batch of that encoding was scanned);
 * the string is valid until the calling thread's next call. For logs and the roofline report. */
const char* mk_engine_scan_kernel(mk_engine* e, mk_encoding enc);

/* Build once, upload N times: the host side of a query set (what AhoCorasick::builder()...build(list) returns in the
 * reference, src/cmd_extract.rs:259-277) as its own object, from which one engine per GPU is created. The seed
 * tables are built once, on host threads, while the callers create their CUDA contexts; every engine uploads its
 * own device copy. mk_engine_create(patterns, config) == mk_tables_create + mk_engine_create_shared +
 * mk_tables_destroy. The engines keep the tables alive: mk_tables_destroy may be called right after the last
 * mk_engine_create_shared. Engines of one mk_tables may be created from different threads at the same time. */
typedef struct mk_tables mk_tables;
int mk_tables_create(const mk_patterns* patterns, int case_insensitive, mk_tables** out);
int mk_engine_create_shared(mk_tables* tables, const mk_config* config, mk_engine** out);
void mk_tables_destroy(mk_tables* t);

/* Slot staging (pinned host memory the caller fills in place). lens_pinned may be NULL if the
 * caller never passes explicit lengths (that buffer is allocated by the first call that asks for it). */
int mk_slot_buffers(mk_engine* e, uint32_t slot, uint8_t** seq_pinned, uint64_t** off_pinned,
                    uint32_t** lens_pinned);
/* Asynchronous: H2D copy of the slot's buffers, scan, hit sort, D2H of the results.
 * n_units = bytes (ASCII) or bases (BAM4) used in seq_pinned; off_pinned holds n_records+1 entries.
 * use_lens != 0: lens_pinned holds n_records record lengths (needed for BAM4 with odd lengths). */
int mk_scan_submit(mk_engine* e, uint32_t slot, uint32_t n_records, uint64_t n_units, int use_lens,
                   mk_encoding enc, mk_mode mode);
int mk_scan_wait(mk_engine* e, uint32_t slot, mk_result* out);

/* Same as submit+wait but H2D-copies from caller-owned host memory (pin it with cudaHostRegister /
 * cudaHostAlloc for full PCIe speed) instead of the slot's pinned buffers. h_lens may be NULL. */
int mk_scan_host(mk_engine* e, uint32_t slot, const uint8_t* h_seq, const uint64_t* h_off,
                 const uint32_t* h_lens, uint32_t n_records, uint64_t n_units, mk_encoding enc,
                 mk_mode mode);

/* mk_scan_host for records of one common length (fixed-length reads): record r occupies units
 * [r * record_len, (r + 1) * record_len) of h_seq, so no offset array crosses the bus — the offsets are
 * written on the device. For MK_ENC_BAM4 record_len must be even. */
int mk_scan_host_uniform(mk_engine* e, uint32_t slot, const uint8_t* h_seq, uint32_t n_records,
                         uint32_t record_len, mk_encoding enc, mk_mode mode);

/* Synchronous scan of a batch that already sits in device memory (sequence decoded / generated on
 * the device, or the device-timed benchmark). d_seq must be 16-byte aligned and readable up to the
 * next multiple of 16 bytes plus 16. fetch != 0 also copies flags and hits to the host views. */
int mk_scan_device(mk_engine* e, const void* d_seq, const uint64_t* d_off, const uint32_t* d_lens,
                   uint32_t n_records, uint64_t n_units, mk_encoding enc, mk_mode mode, int fetch,
                   mk_result* out);

/* Asynchronous flavour of mk_scan_device: enqueue the batch on `slot`'s stream and return; mk_scan_wait
 * on the same slot delivers the result. Lets a caller whose data already sits in device memory keep
 * several batches in flight (the device-timed benchmark; a decoder that produces batches on the GPU). */
int mk_scan_device_submit(mk_engine* e, uint32_t slot, const void* d_seq, const uint64_t* d_off,
                          const uint32_t* d_lens, uint32_t n_records, uint64_t n_units,
                          mk_encoding enc, mk_mode mode, int fetch);

const char* mk_last_error(void);
const char* mk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MERKURIO_CUDA_H */
