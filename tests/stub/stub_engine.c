/* Test double of the C ABI (include/merkurio_cuda.h) for the CPU tier of the test-suite: it owns slot
 * buffers in ordinary memory and reports NO hits for every batch. It is not a matcher and is never
 * shipped or loaded by the product; tests preload it (LD_PRELOAD) to exercise the host's reader ->
 * packer -> driver -> writer plumbing (batching, pieces of long records, chunk hand-over, error
 * propagation, output formatting) on a machine without a GPU.
 *
 * MK_STUB_FLAG_FIRST=<base> (A, C, G or T) makes it pretend instead: every record that starts with that
 * base and is at least as long as the first query gets its flag bit and, outside FLAG mode, one made-up
 * "hit" of the first query at position 0. Nothing is compared with the queries — the point is only that
 * the host's paths (which cut their batches differently) must then agree on which records are written,
 * tagged and logged, flagged and unflagged ones mixed. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "merkurio_cuda.h"

struct mk_engine {
    mk_config cfg;
    uint8_t* seq[16];
    uint64_t* off[16];
    uint32_t* lens[16];
    uint64_t* flags;
    uint32_t n_records[16];
    uint64_t n_units[16];
    int use_lens[16];
    mk_encoding enc[16];
    mk_mode mode[16];
    uint32_t first_len;  /* length of query 0 */
    mk_hit* hits;
};

int mk_engine_create(const mk_patterns* p, const mk_config* c, mk_engine** out) {
    (void)p;
    /* MK_STUB_STARTUP_MS: the time a CUDA context takes to come up (scripts/bench_host_stub.py: the readers run ahead
       meanwhile, as they do in front of the real engine) */
    const char* ms = getenv("MK_STUB_STARTUP_MS");
    if (ms && *ms) usleep((useconds_t)atoi(ms) * 1000);
    mk_engine* e = calloc(1, sizeof *e);
    e->cfg = *c;
    for (uint32_t s = 0; s < c->n_slots && s < 16; ++s) {
        e->seq[s] = calloc(c->max_batch_bytes + 64, 1);
        e->off[s] = calloc((size_t)c->max_batch_records + 1, 8);
        e->lens[s] = calloc((size_t)c->max_batch_records + 1, 4);
    }
    e->flags = calloc((size_t)c->max_batch_records / 64 + 2, 8);
    e->hits = calloc((size_t)c->max_batch_records + 1, sizeof(mk_hit));
    e->first_len = p && p->n ? p->off[1] - p->off[0] : 0;
    *out = e;
    return 0;
}
/* shared tables: the stub only remembers the length of query 0 */
struct mk_tables { uint32_t first_len; int case_insensitive; };
int mk_tables_create(const mk_patterns* p, int ci, mk_tables** out) {
    mk_tables* t = calloc(1, sizeof *t);
    t->first_len = p && p->n ? p->off[1] - p->off[0] : 0;
    t->case_insensitive = ci;
    *out = t;
    return 0;
}
void mk_tables_destroy(mk_tables* t) { free(t); }
int mk_engine_create_shared(mk_tables* t, const mk_config* c, mk_engine** out) {
    int rc = mk_engine_create(NULL, c, out);
    if (rc == 0) (*out)->first_len = t->first_len;
    return rc;
}
const char* mk_engine_scan_kernel(mk_engine* e, mk_encoding enc) { (void)e; (void)enc; return "stub"; }
void mk_engine_destroy(mk_engine* e) {
    if (!e) return;
    for (int s = 0; s < 16; ++s) { free(e->seq[s]); free(e->off[s]); free(e->lens[s]); }
    free(e->flags);
    free(e->hits);
    free(e);
}
int mk_engine_get_info(mk_engine* e, mk_engine_info* o) { (void)e; memset(o, 0, sizeof *o); return 0; }
int mk_slot_buffers(mk_engine* e, uint32_t s, uint8_t** a, uint64_t** b, uint32_t** c) {
    if (a) *a = e->seq[s];
    if (b) *b = e->off[s];
    if (c) *c = e->lens[s];
    return 0;
}
int mk_scan_submit(mk_engine* e, uint32_t s, uint32_t n, uint64_t u, int l, mk_encoding enc, mk_mode m) {
    e->n_records[s] = n;
    e->n_units[s] = u;
    e->use_lens[s] = l;
    e->enc[s] = enc;
    e->mode[s] = m;
    return 0;
}
/* MK_STUB_DIGEST_FILE: after every batch, "<records> <bases> <digest>\n" of everything scanned so far is written there;
   the digest is the sum over the records (ASCII batches) of an FNV-1a hash of their bases, so it depends on what the
   packer put into the slots, not on how the records were dealt into batches. FASTA pieces that repeat bases change
   it: meant for inputs whose records are not cut. */
static uint64_t g_records, g_bases, g_digest;
static void stub_digest(mk_engine* e, uint32_t s) {
    const char* path = getenv("MK_STUB_DIGEST_FILE");
    if (!path || !*path || e->enc[s] != MK_ENC_ASCII) return;
    for (uint32_t i = 0; i < e->n_records[s]; ++i) {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t p = e->off[s][i]; p < e->off[s][i + 1]; ++p) h = (h ^ e->seq[s][p]) * 1099511628211ull;
        g_digest += h;
        g_bases += e->off[s][i + 1] - e->off[s][i];
    }
    g_records += e->n_records[s];
    FILE* f = fopen(path, "w");
    if (f) {
        fprintf(f, "%llu %llu %016llx\n", (unsigned long long)g_records, (unsigned long long)g_bases, (unsigned long long)g_digest);
        fclose(f);
    }
}

int mk_scan_wait(mk_engine* e, uint32_t s, mk_result* r) {
    memset(r, 0, sizeof *r);
    stub_digest(e, s);
    r->record_flags = e->flags;  /* all zero: no record has a hit */
    r->n_records = e->n_records[s];
    r->bases_scanned = e->n_units[s];
    memset(e->flags, 0, ((size_t)e->cfg.max_batch_records / 64 + 2) * 8);
    const char* want = getenv("MK_STUB_FLAG_FIRST");
    if (want && *want) {
        const uint8_t nib = *want == 'A' ? 1 : *want == 'C' ? 2 : *want == 'G' ? 4 : 8;
        uint64_t nh = 0;
        for (uint32_t i = 0; i < e->n_records[s]; ++i) {
            const uint64_t at = e->off[s][i];
            const uint64_t len = e->use_lens[s] ? e->lens[s][i] : e->off[s][i + 1] - at;
            if (len == 0 || len < e->first_len) continue;
            const int starts = e->enc[s] == MK_ENC_ASCII ? e->seq[s][at] == (uint8_t)*want : (e->seq[s][at / 2] >> 4) == nib;
            if (!starts) continue;
            e->flags[i >> 6] |= (uint64_t)1 << (i & 63);
            if (e->mode[s] != MK_MODE_FLAG) {
                mk_hit h = {i, 0, 0, e->first_len};
                e->hits[nh++] = h;
            }
        }
        r->hits = nh ? e->hits : NULL;
        r->n_hits = nh;
    }
    return 0;
}
int mk_scan_host(mk_engine* e, uint32_t s, const uint8_t* a, const uint64_t* b, const uint32_t* c, uint32_t n, uint64_t u,
                 mk_encoding enc, mk_mode m) {
    (void)a; (void)b; (void)c;
    return mk_scan_submit(e, s, n, u, 0, enc, m);
}
int mk_scan_host_uniform(mk_engine* e, uint32_t s, const uint8_t* a, uint32_t n, uint32_t len, mk_encoding enc, mk_mode m) {
    (void)a;
    return mk_scan_submit(e, s, n, (uint64_t)n * len, 0, enc, m);
}
int mk_scan_device(mk_engine* e, const void* a, const uint64_t* b, const uint32_t* c, uint32_t n, uint64_t u, mk_encoding enc,
                   mk_mode m, int f, mk_result* r) {
    (void)e; (void)a; (void)b; (void)c; (void)n; (void)u; (void)enc; (void)m; (void)f;
    memset(r, 0, sizeof *r);
    return 0;
}
int mk_scan_device_submit(mk_engine* e, uint32_t s, const void* a, const uint64_t* b, const uint32_t* c, uint32_t n, uint64_t u,
                          mk_encoding enc, mk_mode m, int f) {
    (void)a; (void)b; (void)c; (void)f;
    return mk_scan_submit(e, s, n, u, 0, enc, m);
}
const char* mk_last_error(void) { return ""; }
const char* mk_version(void) { return "stub (reports no hits; tests only)"; }
