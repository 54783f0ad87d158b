/* mk_oracle.c — CPU restatement of MerKurio's matching path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. The product (merkurio_b200/, include/) never links or calls it.
 *
 * The reference is Rust and cannot be built in this image (no cargo/rustc), so this file restates
 * its algorithms in plain C, citing the lines it follows (paths relative to the reference
 * repository):
 *   - generate_masks            src/pattern_preprocessing.rs:24-43
 *   - BNDMq::new / Matches::next / find_matches / tune_q_value
 *                               src/pattern_matching.rs:61-78, 165-209, 82-125, 213-225
 *   - Aho-Corasick DFA, MatchKind::Standard, find_overlapping_iter, ascii_case_insensitive:
 *     third-party crate aho-corasick 1.1.3 (Cargo.lock:12-13), NOT in the reference tree. Its
 *     published algorithm is restated: trie in pattern order, BFS failure links, a state's match
 *     list = its own patterns followed by the match list of its failure state, DFA over byte
 *     classes with premultiplied state ids, overlapping search reporting every pattern of every
 *     visited match state in list order. Call sites: src/cmd_extract.rs:260-265,332,480,507;
 *     src/cmd_tag.rs:235-240,393-396.
 * Pinned by the reference's own vectors: tests/oracle tests replay src/pattern_matching.rs:353-392,
 * 467-482, src/pattern_preprocessing.rs:54-84 and every golden under tests/fixtures and
 * example-workflow (see tests/golden/). The case-insensitive order of patterns that are equal up
 * to case is not covered by any reference vector ("parity unpinned" for that detail only).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* BNDMq                                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* src/pattern_preprocessing.rs:24-43. Returns 0, or -1 for PatternTooLong (m > 64). */
int mko_generate_masks(const uint8_t* pattern, size_t m, uint64_t masks[256], uint64_t* accept) {
    if (m > 64) return -1;
    memset(masks, 0, 256 * sizeof(uint64_t));
    for (size_t j = 0; j < m; ++j) masks[pattern[j]] |= (uint64_t)1 << (m - j - 1);
    *accept = m ? (uint64_t)1 << (m - 1) : 0;
    return 0;
}

/* src/pattern_matching.rs:213-225. Returns 0 for lengths > 64 (the reference bails). */
int mko_tune_q_value(size_t pattern_len) {
    if (pattern_len <= 1) return 1;
    if (pattern_len <= 3) return 2;
    if (pattern_len <= 8) return 3;
    if (pattern_len <= 30) return 4;
    if (pattern_len <= 55) return 5;
    if (pattern_len <= 64) return 6;
    return 0;
}

/* src/pattern_matching.rs:61-78 error checks: -2 EmptyPattern, -3 InvalidQGramLength, -1 PatternTooLong */
int mko_bndmq_check(size_t m, size_t q) {
    if (m == 0) return -2;
    if (q == 0 || q > m) return -3;
    if (m > 64) return -1;
    return 0;
}

/* Matches::next, src/pattern_matching.rs:165-209, run to exhaustion (== find_all, :151-153).
 * Writes up to cap start offsets, returns the total number of matches (or <0 on a bad pattern).
 * stop_at_first != 0 gives find_match (:128-130): returns 1 as soon as one match exists. */
long mko_bndmq_find(const uint8_t* pattern, size_t m, size_t q, const uint8_t* text, size_t n, uint64_t* out, size_t cap,
                    int stop_at_first) {
    int rc = mko_bndmq_check(m, q);
    if (rc) return rc;
    uint64_t masks[256], accept;
    mko_generate_masks(pattern, m, masks, &accept);
    if (m > n) return 0;
    long found = 0;
    const size_t step = m - q + 1;
    size_t i = step;
    while (i <= n - q + 1) {
        uint64_t d = masks[text[i - 1]];
        for (size_t ii = 0; ii + 1 < q; ++ii) d &= masks[text[i + ii]] << (ii + 1);
        if (d != 0) {
            size_t j = i;
            const size_t first = i - step;
            for (;;) {
                --j;
                if (d >= accept) {
                    if (j > first) {
                        i = j;
                    } else {
                        if (stop_at_first) return 1;
                        if ((size_t)found < cap) out[found] = j;
                        ++found;
                        /* the iterator advances i and returns; resuming re-enters the outer loop */
                        goto advance;
                    }
                }
                /* j == 0 can only be reached right after a reported match (j == first == 0) */
                d = (d << 1) & masks[text[j - 1]];
                if (d == 0) break;
            }
        }
    advance:
        i += step;
    }
    return found;
}

/* all start offsets s with text[s..s+m) == pattern */
long mko_naive_find(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n, uint64_t* out, size_t cap) {
    long found = 0;
    if (m == 0 || m > n) return 0;
    for (size_t s = 0; s + m <= n; ++s)
        if (memcmp(text + s, pattern, m) == 0) {
            if ((size_t)found < cap) out[found] = s;
            ++found;
        }
    return found;
}

/* ------------------------------------------------------------------------------------------ */
/* Aho-Corasick DFA                                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t n_states;     /* DFA states, ids 0..n_states-1; premultiplied by stride in trans */
    uint32_t stride;       /* power of two >= number of byte classes */
    uint32_t n_match;      /* states 1..n_match are the match states (0 is the start state unless it matches) */
    uint32_t start;        /* premultiplied start state id */
    uint8_t classes[256];  /* byte -> class */
    uint32_t* trans;       /* n_states * stride premultiplied next ids */
    uint32_t* match_off;   /* n_states + 1 offsets into match_pat, indexed by state id */
    uint32_t* match_pat;   /* pattern ids */
    uint32_t* pat_len;
    uint32_t n_patterns;
} mko_ac;

static inline uint8_t fold_ascii(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c | 0x20) : c; }

void mko_ac_free(mko_ac* ac) {
    if (!ac) return;
    free(ac->trans); free(ac->match_off); free(ac->match_pat); free(ac->pat_len); free(ac);
}

mko_ac* mko_ac_build(const uint8_t* bytes, const uint32_t* off, uint32_t n, int case_insensitive) {
    mko_ac* ac = calloc(1, sizeof *ac);
    if (!ac) return NULL;
    /* byte classes: one class per distinct (folded) pattern byte, class 0 = every other byte */
    uint32_t ncls = 1;
    int cls_of[256];
    for (int b = 0; b < 256; ++b) cls_of[b] = -1;
    for (uint32_t i = off[0]; i < off[n]; ++i) {
        uint8_t c = case_insensitive ? fold_ascii(bytes[i]) : bytes[i];
        if (cls_of[c] < 0) cls_of[c] = (int)ncls++;
    }
    for (int b = 0; b < 256; ++b) {
        uint8_t c = case_insensitive ? fold_ascii((uint8_t)b) : (uint8_t)b;
        ac->classes[b] = (uint8_t)(cls_of[c] < 0 ? 0 : cls_of[c]);
    }
    uint32_t stride = 1;
    while (stride < ncls) stride <<= 1;

    /* trie, patterns inserted in list order */
    size_t total = (size_t)(off[n] - off[0]) + 1;
    uint32_t* next = malloc(total * ncls * sizeof(uint32_t));
    uint32_t* own_first = malloc(total * sizeof(uint32_t));  /* head of this state's own pattern chain */
    uint32_t* own_next = malloc((size_t)n * sizeof(uint32_t));
    uint32_t* own_tail = malloc(total * sizeof(uint32_t));
    uint32_t* fail = calloc(total, sizeof(uint32_t));
    ac->pat_len = malloc((size_t)n * sizeof(uint32_t));
    if (!next || !own_first || !own_next || !own_tail || !fail || !ac->pat_len) return NULL;
    const uint32_t NONE = 0xFFFFFFFFu;
    uint32_t ns = 1;
    for (uint32_t c = 0; c < ncls; ++c) next[c] = NONE;
    own_first[0] = own_tail[0] = NONE;
    for (uint32_t p = 0; p < n; ++p) {
        uint32_t s = 0;
        ac->pat_len[p] = off[p + 1] - off[p];
        for (uint32_t i = off[p]; i < off[p + 1]; ++i) {
            uint32_t c = ac->classes[bytes[i]];
            if (next[(size_t)s * ncls + c] == NONE) {
                for (uint32_t k = 0; k < ncls; ++k) next[(size_t)ns * ncls + k] = NONE;
                own_first[ns] = own_tail[ns] = NONE;
                next[(size_t)s * ncls + c] = ns++;
            }
            s = next[(size_t)s * ncls + c];
        }
        own_next[p] = NONE;
        if (own_first[s] == NONE) own_first[s] = p; else own_next[own_tail[s]] = p;
        own_tail[s] = p;
    }
    /* BFS order + failure links; missing transitions are filled in from the failure state, which
       turns the trie into the DFA (the start state loops on itself for unknown bytes) */
    uint32_t* order = malloc((size_t)ns * sizeof(uint32_t));
    uint32_t* n_matches = calloc(ns, sizeof(uint32_t));
    if (!order || !n_matches) return NULL;
    uint32_t qh = 0, qt = 0;
    for (uint32_t c = 0; c < ncls; ++c) {
        uint32_t t = next[c];
        if (t == NONE) next[c] = 0; else { fail[t] = 0; order[qt++] = t; }
    }
    while (qh < qt) {
        uint32_t s = order[qh++];
        for (uint32_t c = 0; c < ncls; ++c) {
            uint32_t t = next[(size_t)s * ncls + c];
            uint32_t via_fail = next[(size_t)fail[s] * ncls + c];
            if (t == NONE) next[(size_t)s * ncls + c] = via_fail;
            else { fail[t] = via_fail; order[qt++] = t; }
        }
    }
    /* match lists: own patterns first, then the failure state's list (already final in BFS order) */
    uint64_t total_matches = 0;
    for (uint32_t s = 0; s < ns; ++s) for (uint32_t p = own_first[s]; p != NONE; p = own_next[p]) n_matches[s]++;
    for (uint32_t k = 0; k < qt; ++k) { uint32_t s = order[k]; n_matches[s] += n_matches[fail[s]]; }
    for (uint32_t s = 0; s < ns; ++s) total_matches += n_matches[s];
    /* renumber: match states get the ids 1..n_match, so "is match" is one compare in the search loop */
    uint32_t* newid = malloc((size_t)ns * sizeof(uint32_t));
    if (!newid) return NULL;
    uint32_t nm = 0, id = 1;
    for (uint32_t s = 1; s < ns; ++s) if (n_matches[s]) { newid[s] = id++; nm++; }
    for (uint32_t s = 1; s < ns; ++s) if (!n_matches[s]) newid[s] = id++;
    newid[0] = 0;
    ac->n_states = ns; ac->stride = stride; ac->n_match = nm; ac->start = 0; ac->n_patterns = n;
    ac->trans = malloc((size_t)ns * stride * sizeof(uint32_t));
    ac->match_off = malloc(((size_t)ns + 1) * sizeof(uint32_t));
    ac->match_pat = malloc((size_t)(total_matches ? total_matches : 1) * sizeof(uint32_t));
    if (!ac->trans || !ac->match_off || !ac->match_pat) return NULL;
    for (uint32_t s = 0; s < ns; ++s)
        for (uint32_t c = 0; c < stride; ++c)
            ac->trans[(size_t)newid[s] * stride + c] = newid[next[(size_t)s * ncls + (c < ncls ? c : 0)]] * stride;
    /* lay the lists out by new id */
    uint32_t* old_of = malloc((size_t)ns * sizeof(uint32_t));
    if (!old_of) return NULL;
    for (uint32_t s = 0; s < ns; ++s) old_of[newid[s]] = s;
    uint32_t pos = 0;
    for (uint32_t i = 0; i < ns; ++i) { ac->match_off[i] = pos; pos += n_matches[old_of[i]]; }
    ac->match_off[ns] = pos;
    /* fill in BFS order so that a failure state's list exists before it is copied */
    {
        uint32_t s = 0, w = ac->match_off[newid[s]];
        for (uint32_t p = own_first[s]; p != NONE; p = own_next[p]) ac->match_pat[w++] = p;
    }
    for (uint32_t k = 0; k < qt; ++k) {
        uint32_t s = order[k], w = ac->match_off[newid[s]];
        for (uint32_t p = own_first[s]; p != NONE; p = own_next[p]) ac->match_pat[w++] = p;
        uint32_t f = newid[fail[s]];
        for (uint32_t r = ac->match_off[f]; r < ac->match_off[f + 1]; ++r) ac->match_pat[w++] = ac->match_pat[r];
    }
    free(next); free(own_first); free(own_next); free(own_tail); free(fail); free(order); free(n_matches); free(newid); free(old_of);
    return ac;
}

uint32_t mko_ac_n_states(const mko_ac* ac) { return ac->n_states; }
uint64_t mko_ac_table_bytes(const mko_ac* ac) { return (uint64_t)ac->n_states * ac->stride * 4; }

/* find_overlapping_iter: every (pattern, start) in report order. Returns the total count; writes
 * at most cap entries. stop_at_first: return 1 at the first match (the `break` at
 * src/cmd_extract.rs:333-335). */
long mko_ac_find_overlapping(const mko_ac* ac, const uint8_t* text, size_t n, uint32_t* out_pat, uint64_t* out_start,
                             size_t cap, int stop_at_first) {
    long found = 0;
    const uint32_t stride = ac->stride, lim = ac->n_match * stride;
    uint32_t sid = ac->start;
    for (size_t i = 0; i < n; ++i) {
        sid = ac->trans[sid + ac->classes[text[i]]];
        if (sid != 0 && sid <= lim) { /* 1*stride <= sid <= n_match*stride */
            if (stop_at_first) return 1;
            uint32_t s = sid / stride;
            for (uint32_t r = ac->match_off[s]; r < ac->match_off[s + 1]; ++r) {
                uint32_t p = ac->match_pat[r];
                if ((size_t)found < cap) { out_pat[found] = p; out_start[found] = i + 1 - ac->pat_len[p]; }
                ++found;
            }
        }
    }
    return found;
}

/* ------------------------------------------------------------------------------------------ */
/* Batch scans (CPU baseline): records = seq[off[r] .. off[r+1])                              */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const mko_ac* ac;
    const uint8_t* seq;
    const uint64_t* off;
    uint32_t r0, r1;
    int count_all;       /* 0: any-hit with early exit (extract without log); 1: count every hit */
    uint64_t* flags;     /* bitmap shared between threads; ranges are 64-record aligned */
    uint64_t n_hits, n_records_hit;
} scan_job;

static void* scan_worker(void* arg) {
    scan_job* j = arg;
    const mko_ac* ac = j->ac;
    const uint32_t stride = ac->stride, hi = ac->n_match * stride;
    for (uint32_t r = j->r0; r < j->r1; ++r) {
        const uint8_t* t = j->seq + j->off[r];
        size_t n = (size_t)(j->off[r + 1] - j->off[r]);
        uint32_t sid = ac->start;
        uint64_t hits = 0;
        for (size_t i = 0; i < n; ++i) {
            sid = ac->trans[sid + ac->classes[t[i]]];
            if (sid != 0 && sid <= hi) {
                uint32_t s = sid / stride;
                hits += ac->match_off[s + 1] - ac->match_off[s];
                if (!j->count_all) break;
            }
        }
        if (hits) {
            j->flags[r >> 6] |= (uint64_t)1 << (r & 63);
            j->n_hits += hits;
            j->n_records_hit++;
        }
    }
    return NULL;
}

/* Scans n_records records with n_threads threads. flags: (n_records+63)/64 words, zeroed here.
 * Returns the number of records with a hit; *n_hits = total hits (count_all) or records hit. */
uint64_t mko_ac_scan_batch(const mko_ac* ac, const uint8_t* seq, const uint64_t* off, uint32_t n_records, int n_threads,
                           int count_all, uint64_t* flags, uint64_t* n_hits) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    memset(flags, 0, ((size_t)n_records + 63) / 64 * 8);
    pthread_t th[256];
    scan_job jobs[256];
    uint32_t words = (n_records + 63) / 64, per = (words + (uint32_t)n_threads - 1) / (uint32_t)n_threads;
    int used = 0;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t r0 = (uint64_t)t * per * 64, r1 = r0 + (uint64_t)per * 64;
        if (r0 >= n_records) break;
        if (r1 > n_records) r1 = n_records;
        jobs[t] = (scan_job){ac, seq, off, (uint32_t)r0, (uint32_t)r1, count_all, flags, 0, 0};
        pthread_create(&th[t], NULL, scan_worker, &jobs[t]);
        used++;
    }
    uint64_t rec = 0, hits = 0;
    for (int t = 0; t < used; ++t) {
        pthread_join(th[t], NULL);
        rec += jobs[t].n_records_hit;
        hits += jobs[t].n_hits;
    }
    if (n_hits) *n_hits = hits;
    return rec;
}

/* All hits of a batch in AC report order, for parity checks against the device hit list.
 * Writes up to cap entries of (record, start, pattern); returns the total. */
uint64_t mko_ac_batch_hits(const mko_ac* ac, const uint8_t* seq, const uint64_t* off, const uint32_t* lens,
                           uint32_t n_records, uint32_t* out_rec, uint32_t* out_start, uint32_t* out_pat, uint64_t cap) {
    uint64_t found = 0;
    const uint32_t stride = ac->stride, hi = ac->n_match * stride;
    for (uint32_t r = 0; r < n_records; ++r) {
        const uint8_t* t = seq + off[r];
        size_t n = lens ? lens[r] : (size_t)(off[r + 1] - off[r]);
        uint32_t sid = ac->start;
        for (size_t i = 0; i < n; ++i) {
            sid = ac->trans[sid + ac->classes[t[i]]];
            if (sid != 0 && sid <= hi) {
                uint32_t s = sid / stride;
                for (uint32_t k = ac->match_off[s]; k < ac->match_off[s + 1]; ++k) {
                    uint32_t p = ac->match_pat[k];
                    if (found < cap) { out_rec[found] = r; out_start[found] = (uint32_t)(i + 1 - ac->pat_len[p]); out_pat[found] = p; }
                    ++found;
                }
            }
        }
    }
    return found;
}
