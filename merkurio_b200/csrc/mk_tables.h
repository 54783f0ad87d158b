// Host-side construction of the device query tables (the role src/pattern_preprocessing.rs:24-43
// generate_masks and the AhoCorasick build at src/cmd_extract.rs:259-277 play in the reference).
//
// Input: the sorted unique pattern list (src/helpers.rs:76-133). Output, per text encoding:
//   * seed geometry (q, d): every d-th text base starts a q-base seed; d + q - 1 <= min pattern
//     length, so each occurrence of each pattern contains exactly one grid seed among its first d
//     offsets;
//   * a cuckoo hash table  seed code -> postings  (4-slot buckets of {code, first posting}; one
//     bucket = one 32-byte L2 sector) and the postings (pattern id, seed offset j in the pattern);
//   * the first-level filter bitmap probed for every seed (staged in shared memory when the seed
//     set is small enough to leave it selective, else left L2-resident);
//   * the pattern bytes used by the exact verify step (case-folded under -I).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "mk_codes.h"

namespace mk {

constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;
constexpr uint32_t kBucketSlots = 4;
constexpr uint32_t kMaxPatternId = (1u << 25) - 2;

struct SeedSlot {
    uint32_t code;
    uint32_t first;  // bit 31: key group (1 = 16-base seed of the long group); bit 30: the key has a single
                     // posting, stored right here as pattern_id << 4 | j (saves the verify kernel one
                     // dependent load); else bits 0..29: index of the first posting. kEmptySlot: free slot
};
constexpr uint32_t kGroupBit = 0x80000000u;
constexpr uint32_t kInlineBit = 0x40000000u;

// posting = pattern_id << 5 | j << 1 | last_of_list
inline uint32_t make_posting(uint32_t pid, uint32_t j, bool last) { return (pid << 5) | (j << 1) | (last ? 1u : 0u); }

struct Tables {
    int enc = 0;
    uint32_t q = 0, d = 0;
    uint32_t q2 = 0;       // 16 when patterns of >= long_min_len bases are indexed by 16-base seeds, else 0
    uint32_t long_min_len = 0;
    bool perm = false;  // D == 16: permuted unit packing
    bool win = false;   // D == 8 / 4 with the shared-memory filter: seeds are masked 16-base windows in the permuted packing
    uint32_t win_mask0 = 0xFFFFFFFFu, win_mask1 = 0xFFFFFFFFu;
    bool filter_dual = false;  // L2-resident 64-bit blocked filter probed with both keys (stride < 16)
    bool filter_direct = false;  // mk_scan_short (ASCII, stride 1 / 2): `filter` is a direct bitmap over the first direct_q1 bases of a seed
    uint32_t direct_q1 = 0;
    bool dual_perm = false;    // dual-key flavour for mk_scan_dual8 (ASCII, stride 8): keys in the permuted window layout, 3-bit filter entries
    // Alphabet gate: a text byte b with (b ^ gate_val) & gate_mask != 0 occurs in no pattern (after -I folding),
    // so a seed window that holds one inside its first q bytes cannot belong to a match. 0: no such bit exists.
    uint32_t gate_mask = 0, gate_val = 0;
    uint32_t n_seeds = 0;
    // first-level filter
    uint32_t filter_log2_bits = 0, filter_hashes = 1;
    uint32_t filter_blocks = 0;  // shared-memory flavour: number of 64-bit blocks
    bool filter_in_smem = true;
    bool filter32 = false;  // shared-memory flavour with 32-bit blocks, 3 bits per key (kernels of stride 16 / 8 / 4 with window or unit seeds)
    std::vector<uint32_t> filter;
    // second-level filter: L2-resident bitmap (~64 bits per seed) probed by the candidates the
    // shared-memory filter lets through, before the cuckoo table is touched (empty: not used)
    uint32_t filter2_log2_bits = 0;
    std::vector<uint32_t> filter2;
    // cuckoo table
    uint32_t bucket_mask = 0;
    std::vector<SeedSlot> slots;  // (bucket_mask + 1) * kBucketSlots
    std::vector<uint32_t> postings;
    // verify data
    std::vector<uint8_t> pat_bytes;   // compare form of every pattern (folded under -I / nibble codes for BAM4)
    std::vector<uint32_t> pat_off;    // n + 1
    std::vector<uint8_t> pat_live;    // 0: pattern can never match in this encoding
};

struct PatternSet {
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> off;  // n + 1
    uint32_t n = 0, min_len = 0, max_len = 0;
    bool case_insensitive = false;
    // sort-key geometry (shared by both encodings)
    uint32_t len_bits = 0, tie_bits = 0;
    std::vector<uint32_t> tie_rank;  // rank inside the class of fold-equal patterns
    uint32_t len(uint32_t p) const { return off[p + 1] - off[p]; }
    const uint8_t* ptr(uint32_t p) const { return bytes.data() + off[p]; }
};

inline uint32_t bits_for(uint64_t v) {
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

// Seed geometry from the shortest pattern: the largest stride whose seed is still selective.
inline void choose_geometry(uint32_t min_len, uint32_t* q, uint32_t* d) {
    const uint32_t strides[5] = {16, 8, 4, 2, 1};
    for (uint32_t s : strides) {
        if (min_len < s) continue;
        uint32_t qq = std::min<uint32_t>(16, min_len - s + 1);
        // stride 16 uses the permuted whole-unit packing, which needs the full 16-base seed
        if (s == 16 ? qq == 16 : (qq >= 12 || s == 1)) { *q = qq; *d = s; return; }
    }
    *q = 1; *d = 1;
}

inline PatternSet make_pattern_set(const uint8_t* bytes, const uint32_t* off, uint32_t n, bool case_insensitive) {
    PatternSet ps;
    ps.n = n;
    ps.case_insensitive = case_insensitive;
    ps.off.assign(off, off + n + 1);
    ps.bytes.assign(bytes + off[0], bytes + off[n]);
    for (auto& o : ps.off) o -= off[0];
    ps.min_len = UINT32_MAX;
    for (uint32_t p = 0; p < n; ++p) {
        ps.min_len = std::min(ps.min_len, ps.len(p));
        ps.max_len = std::max(ps.max_len, ps.len(p));
    }
    ps.len_bits = bits_for(ps.max_len);
    // Patterns equal up to ASCII case share every span under -I; they are reported in ascending
    // pattern index (the order aho-corasick 1.1.3 appends pattern ids to a trie state).
    ps.tie_rank.assign(n, 0);
    uint32_t max_rank = 0;
    if (case_insensitive) {
        std::vector<uint32_t> idx(n);
        for (uint32_t i = 0; i < n; ++i) idx[i] = i;
        auto folded_less = [&](uint32_t a, uint32_t b) {
            uint32_t la = ps.len(a), lb = ps.len(b);
            const uint8_t *pa = ps.ptr(a), *pb = ps.ptr(b);
            for (uint32_t i = 0; i < std::min(la, lb); ++i) {
                uint8_t ca = mk_fold(pa[i]), cb = mk_fold(pb[i]);
                if (ca != cb) return ca < cb;
            }
            if (la != lb) return la < lb;
            return a < b;
        };
        auto folded_eq = [&](uint32_t a, uint32_t b) {
            if (ps.len(a) != ps.len(b)) return false;
            for (uint32_t i = 0; i < ps.len(a); ++i)
                if (mk_fold(ps.ptr(a)[i]) != mk_fold(ps.ptr(b)[i])) return false;
            return true;
        };
        std::sort(idx.begin(), idx.end(), folded_less);
        for (uint32_t i = 1; i < n; ++i)
            if (folded_eq(idx[i - 1], idx[i])) {
                ps.tie_rank[idx[i]] = ps.tie_rank[idx[i - 1]] + 1;
                max_rank = std::max(max_rank, ps.tie_rank[idx[i]]);
            }
    }
    ps.tie_bits = bits_for(max_rank);
    return ps;
}

// class of one compare-form symbol (ASCII byte or BAM nibble)
inline uint32_t sym_class(int enc, uint8_t s) {
    if (enc == 0) return (s >> 1) & 3u;
    uint32_t n = s & 0xF;
    return ((((n >> 2) | (n >> 3)) & 1u) << 1) | (((n >> 1) | (n >> 3)) & 1u);
}

// seed code of q symbols starting at sym[0] (ordered packing, as mk_seed_ord yields it)
inline uint32_t seed_code_ord(int enc, const uint8_t* sym, uint32_t q) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < q; ++i) c = (c << 2) | sym_class(enc, sym[i]);
    return c;
}
// seed code of exactly 16 symbols with the permuted packing of the device fast path
inline uint32_t seed_code_perm(int enc, const uint8_t* sym) {
    if (enc == 0) {
        uint32_t w[4];
        std::memcpy(w, sym, 16);
        return mk_pack_ascii_perm(w[0], w[1], w[2], w[3]);
    }
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = (uint8_t)((sym[2 * i] << 4) | (sym[2 * i + 1] & 0xF));
    uint32_t w[2];
    std::memcpy(w, b, 8);
    return mk_pack_bam_perm(w[0], w[1]);
}

struct SeedKey {
    uint32_t code, first, group;
};

// Window layout of mk_scan_win: the q bases of a 16-base window that form the seed, as a mask of the
// permuted ASCII code (base t: byte lane t % 4, 2-bit field t / 4) or of the two BAM4 words (base t:
// byte t / 2 of the 8-byte window, high nibble first).
inline void window_masks(int enc, uint32_t q, uint32_t* m0, uint32_t* m1) {
    *m0 = *m1 = 0;
    for (uint32_t t = 0; t < q && t < 16; ++t) {
        if (enc == 0) {
            *m0 |= 3u << (8 * (t % 4) + 2 * (t / 4));
        } else {
            uint32_t byte = t / 2, bits = 0xFu << (8 * (byte % 4) + (t % 2 == 0 ? 4 : 0));
            if (byte < 4) *m0 |= bits; else *m1 |= bits;
        }
    }
    if (enc == 0) *m1 = 0;
}
// seed code of the first q (<= 16) symbols at sym in the window layout
inline uint32_t seed_code_win(int enc, const uint8_t* sym, uint32_t q, uint32_t m0, uint32_t m1) {
    uint8_t s16[16] = {0};
    for (uint32_t t = 0; t < q && t < 16; ++t) s16[t] = sym[t];
    if (enc == 0) {
        uint32_t w[4];
        std::memcpy(w, s16, 16);
        return mk_pack_ascii_perm(w[0], w[1], w[2], w[3]) & m0;
    }
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = (uint8_t)((s16[2 * i] << 4) | (s16[2 * i + 1] & 0xF));
    uint32_t w[2];
    std::memcpy(w, b, 8);
    return mk_pack_bam_perm(w[0] & m0, w[1] & m1);
}

inline bool cuckoo_build(const std::vector<SeedKey>& keys, uint32_t log2_buckets, std::vector<SeedSlot>* out,
                         uint32_t* mask_out) {
    uint32_t nb = 1u << log2_buckets, mask = nb - 1;
    std::vector<SeedSlot> slots((size_t)nb * kBucketSlots, SeedSlot{0, kEmptySlot});
    std::mt19937 rng(0x5EEDu + log2_buckets);
    const size_t kAhead = 24;  // the table is far larger than the cache: fetch the buckets of the keys to come
    for (size_t ki = 0; ki < keys.size(); ++ki) {
        if (ki + kAhead < keys.size()) {
            const SeedKey& nx = keys[ki + kAhead];
            const uint32_t hk = mk_group_key(nx.code, nx.group ? 1u : 0u);
            __builtin_prefetch(&slots[(size_t)mk_hash_b1(hk, mask) * kBucketSlots], 1);
            __builtin_prefetch(&slots[(size_t)mk_hash_b2(hk, mask) * kBucketSlots], 1);
        }
        const SeedKey& kv = keys[ki];
        SeedSlot cur{kv.code, kv.first | (kv.group ? kGroupBit : 0u)};
        bool placed = false;
        uint32_t from = UINT32_MAX;
        for (int kick = 0; kick < 500 && !placed; ++kick) {
            uint32_t hk = mk_group_key(cur.code, cur.first >> 31);
            uint32_t cand[2] = {mk_hash_b1(hk, mask), mk_hash_b2(hk, mask)};
            for (int h = 0; h < 2 && !placed; ++h)
                for (uint32_t s = 0; s < kBucketSlots && !placed; ++s) {
                    SeedSlot& sl = slots[(size_t)cand[h] * kBucketSlots + s];
                    if (sl.first == kEmptySlot) { sl = cur; placed = true; }
                }
            if (placed) break;
            // evict a random slot, preferring the bucket the item did not just come from
            uint32_t b = (cand[0] == from) ? cand[1] : (cand[1] == from) ? cand[0] : cand[rng() & 1];
            std::swap(cur, slots[(size_t)b * kBucketSlots + rng() % kBucketSlots]);
            from = b;
        }
        if (!placed) return false;
    }
    *out = std::move(slots);
    *mask_out = mask;
    return true;
}

// Expected false-positive rate of one probe of a blocked Bloom filter with `nblocks` 64-bit blocks
// holding nn keys: Poisson(lambda) keys per block, each key sets 2 bits in each 32-bit half.
inline double blocked_fp(double nn, uint32_t nblocks) {
    double lambda = nn / nblocks, fp = 0.0, pmf = std::exp(-lambda);
    for (int j = 0; j < 600; ++j) {
        double set = 1.0 - std::pow(31.0 / 32.0, 2.0 * j);
        fp += pmf * std::pow(set, 4.0);
        pmf *= lambda / (j + 1);
    }
    return fp;
}

// The same for 32-bit blocks with 3 bits per key.
inline double blocked_fp32(double nn, uint32_t nblocks32) {
    double lambda = nn / nblocks32, fp = 0.0, pmf = std::exp(-lambda);
    for (int j = 0; j < 600; ++j) {
        double set = 1.0 - std::pow(31.0 / 32.0, 3.0 * j);
        fp += pmf * std::pow(set, 3.0);
        pmf *= lambda / (j + 1);
    }
    return fp;
}

#ifdef MK_TABLE_TIMING
#define TBT_INIT double tbt_t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
#define TBT(name) { double n_ = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); std::fprintf(stderr, "  %s %.3f\n", name, n_ - tbt_t0); tbt_t0 = n_; }
#else
#define TBT_INIT
#define TBT(name)
#endif
inline Tables build_tables(const PatternSet& ps, int enc) {
    Tables t;
    TBT_INIT
    t.enc = enc;
    const uint32_t n = ps.n;
    // compare form of the patterns
    t.pat_off = ps.off;
    t.pat_bytes.resize(ps.bytes.size());
    t.pat_live.assign(n, 1);
    if (enc == 0 && !ps.case_insensitive && !ps.bytes.empty()) std::memcpy(t.pat_bytes.data(), ps.bytes.data(), ps.bytes.size());
    else for (uint32_t p = 0; p < n; ++p) {
        for (uint32_t i = ps.off[p]; i < ps.off[p + 1]; ++i) {
            uint8_t c = ps.bytes[i];
            if (enc == 0) {
                t.pat_bytes[i] = ps.case_insensitive ? mk_fold(c) : c;
            } else {
                // decoded BAM text is upper-case "=ACMGRSVTWYHKDBN"; under -I lower-case query letters match it
                if (ps.case_insensitive && c >= 'a' && c <= 'z') c = (uint8_t)(c - 0x20);
                uint8_t nib = mk_ascii_to_nibble(c);
                if (nib == 0xFF) { t.pat_live[p] = 0; nib = 0; }
                t.pat_bytes[i] = nib;
            }
        }
    }
    t.pat_bytes.resize(ps.bytes.size() + 16, 0);  // the verify kernel compares with 8-byte loads at any offset
    choose_geometry(ps.min_len, &t.q, &t.d);
    t.perm = (t.d == 16);

    // Every indexed seed as one 64-bit word: group (bit 63) | code (bits 31..62) | pattern (bits 4..30) | offset j
    // (bits 0..3), sorted — i.e. by (group, code, pattern, j) — and the distinct keys with their first posting.
    // The words are generated in (pattern, j) order, so a stable LSD radix sort over the 33 key bits is enough
    // (8-bit digits; std::sort of 8 M seeds was most of the 7 s a million queries took to index).
    std::vector<uint64_t> trips, trips_tmp;
    std::vector<SeedKey> keys;
    auto t_grp = [](uint64_t v) { return (uint32_t)(v >> 63); };
    auto t_code = [](uint64_t v) { return (uint32_t)(v >> 31); };
    auto t_pid = [](uint64_t v) { return (uint32_t)(v >> 4) & 0x7FFFFFFu; };
    auto t_j = [](uint64_t v) { return (uint32_t)v & 15u; };
    // stride 8 / 4 (BAM4: 8 only) can use the window layout of mk_scan_win if the filter fits shared memory
    bool win_layout = (enc == 0 ? (t.d == 8 || t.d == 4) : t.d == 8) && !std::getenv("MK_NO_WIN_SCAN");
    if (win_layout) window_masks(enc, t.q, &t.win_mask0, &t.win_mask1);
    // ordered code byte b (bases 4b .. 4b+3, first base in the two highest bits) -> its fields in the permuted
    // packing of mk_pack_ascii_perm (base t: byte lane t % 4, 2-bit field t / 4)
    uint32_t perm_lut[4][256];
    for (uint32_t b = 0; b < 4; ++b)
        for (uint32_t v = 0; v < 256; ++v) {
            uint32_t w = 0;
            for (uint32_t i = 0; i < 4; ++i) w |= ((v >> (6 - 2 * i)) & 3u) << (8 * i + 2 * b);
            perm_lut[b][v] = w;
        }
    // the seeds of patterns [p0, p1), d words per live pattern, written to out
    auto seeds_of = [&](uint32_t p0, uint32_t p1, uint32_t long_min_len, uint64_t* out) {
        for (uint32_t p = p0; p < p1; ++p) {
            if (!t.pat_live[p]) continue;
            const uint8_t* sym = t.pat_bytes.data() + t.pat_off[p];
            const bool lng = long_min_len && ps.len(p) >= long_min_len;
            const uint64_t tag = (uint64_t)(lng ? 1u : 0u) << 63 | (uint64_t)p << 4;
            if (enc == 0) {
                // ASCII: every layout is a function of the 2-bit classes, so the d seeds of a pattern come from
                // one rolling code (ordered packing) and, for the permuted / window layouts, four table look-ups
                const uint32_t qq = t.perm ? 16u : win_layout ? std::min(t.q, 16u) : (lng ? 16u : t.q);
                const uint32_t keep = qq == 16 ? 0xFFFFFFFFu : (1u << (2 * qq)) - 1u;
                uint32_t c = 0;
                for (uint32_t i = 0; i + 1 < qq; ++i) c = (c << 2) | sym_class(0, sym[i]);
                for (uint32_t j = 0; j < t.d; ++j) {
                    c = ((c << 2) | sym_class(0, sym[j + qq - 1])) & keep;
                    uint32_t code = c;
                    if (t.perm || win_layout || t.dual_perm) {
                        const uint32_t top = c << (32 - 2 * qq);  // first base in the two highest bits
                        code = perm_lut[0][top >> 24] | perm_lut[1][(top >> 16) & 255] | perm_lut[2][(top >> 8) & 255] | perm_lut[3][top & 255];
                    }
                    *out++ = tag | (uint64_t)code << 31 | j;
                }
                continue;
            }
            for (uint32_t j = 0; j < t.d; ++j) {
                uint32_t code = t.perm ? seed_code_perm(enc, sym + j)
                              : win_layout ? seed_code_win(enc, sym + j, t.q, t.win_mask0, t.win_mask1)
                                           : seed_code_ord(enc, sym + j, lng ? 16u : t.q);
                *out++ = tag | (uint64_t)code << 31 | j;
            }
        }
    };
    size_t n_live = 0;
    for (uint32_t p = 0; p < n; ++p) n_live += t.pat_live[p] ? 1 : 0;
    // large query sets: a few threads generate and sort (both encodings are usually built at the same time)
    const size_t n_threads = n_live * t.d < ((size_t)1 << 20) ? 1 : std::min<size_t>(4, std::max(1u, std::thread::hardware_concurrency() / 2));
    auto on_threads = [&](const std::function<void(size_t)>& job) {
        if (n_threads == 1) return job(0);
        std::vector<std::thread> th;
        for (size_t w = 0; w < n_threads; ++w) th.emplace_back(job, w);
        for (auto& x : th) x.join();
    };
    auto index_seeds = [&](uint32_t long_min_len) {
        keys.clear();
        if (n_live * t.d >= (size_t)kInlineBit) throw std::runtime_error("too many seeds for the posting index");
        trips.resize(n_live * t.d);
        {
            // pattern ranges of equal size; where each range's words start follows from the live patterns before it
            std::vector<uint32_t> cut(n_threads + 1);
            std::vector<size_t> first(n_threads + 1, 0);
            for (size_t w = 0; w <= n_threads; ++w) cut[w] = (uint32_t)((uint64_t)n * w / n_threads);
            for (size_t w = 0; w < n_threads; ++w) {
                size_t live = 0;
                for (uint32_t p = cut[w]; p < cut[w + 1]; ++p) live += t.pat_live[p] ? 1 : 0;
                first[w + 1] = first[w] + live * t.d;
            }
            on_threads([&](size_t w) { seeds_of(cut[w], cut[w + 1], long_min_len, trips.data() + first[w]); });
        }
        if (trips.size() < 4096) {
            std::sort(trips.begin(), trips.end());
        } else {
            trips_tmp.resize(trips.size());
            // 8-bit digits: 256 write streams stay in the cache (2048 did not); a pass whose digit is the same in
            // every word (the group bit of a single-group index, the empty top bits of short seeds) is skipped.
            // Each thread counts and scatters one contiguous slice; slices keep their order inside a digit (stable).
            const size_t total = trips.size();
            std::vector<size_t> count(n_threads * 256);
            for (int shift = 31; shift < 64; shift += 8) {
                const uint64_t* src = trips.data();
                uint64_t* dst = trips_tmp.data();
                on_threads([&](size_t w) {
                    size_t* c = count.data() + w * 256;
                    std::fill(c, c + 256, (size_t)0);
                    for (size_t i = total * w / n_threads, e = total * (w + 1) / n_threads; i < e; ++i) ++c[(src[i] >> shift) & 255];
                });
                bool one_digit = false;
                size_t at = 0;
                for (size_t dg = 0; dg < 256; ++dg) {
                    size_t of_digit = 0;
                    for (size_t w = 0; w < n_threads; ++w) {
                        const size_t k = count[w * 256 + dg];
                        count[w * 256 + dg] = at;
                        at += k;
                        of_digit += k;
                    }
                    one_digit |= (of_digit == total);
                }
                if (one_digit) continue;
                on_threads([&](size_t w) {
                    size_t* c = count.data() + w * 256;
                    for (size_t i = total * w / n_threads, e = total * (w + 1) / n_threads; i < e; ++i) dst[c[(src[i] >> shift) & 255]++] = src[i];
                });
                trips.swap(trips_tmp);
            }
        }
        const uint64_t key_bits = ~(uint64_t)0 << 31;
        keys.reserve(trips.size());
        for (size_t i = 0; i < trips.size(); ++i)
            if (i == 0 || ((trips[i - 1] ^ trips[i]) & key_bits)) {
                const bool single = i + 1 == trips.size() || ((trips[i] ^ trips[i + 1]) & key_bits);
                keys.push_back({t_code(trips[i]), single ? (kInlineBit | (t_pid(trips[i]) << 4) | t_j(trips[i])) : (uint32_t)i, t_grp(trips[i])});
            }
    };
    TBT("pat_bytes");
    index_seeds(0);
    TBT("index1");

    // first-level filter flavour: blocked Bloom in shared memory while it stays selective
    uint32_t nblocks = MK_BLOOM_MIN_BLOCKS;
    if (blocked_fp((double)keys.size(), nblocks) > 0.002) nblocks = MK_BLOOM_MAX_BLOCKS;
    if (const char* fb = std::getenv("MK_FILTER_BLOCKS")) {  // tuning override
        uint32_t v = (uint32_t)std::atoi(fb) & ~1u;
        if (v >= 1024 && v <= MK_BLOOM_MAX_BLOCKS) nblocks = v;
    }
    // MK_FILTER_MODE=l2|smem overrides the choice (tests exercise both paths on small inputs)
    const char* force = std::getenv("MK_FILTER_MODE");
    bool want_smem = blocked_fp((double)keys.size(), nblocks) <= 0.25;
    if (force && std::strcmp(force, "l2") == 0) want_smem = false;
    if (force && std::strcmp(force, "smem") == 0) want_smem = true;
    // Seed sets too large for shared memory with a stride below 16 are indexed once more below, with 16-base seeds
    // for the patterns that are long enough
    const bool long_seeds = !want_smem && t.d < 16 && t.d >= 2 && t.q < 16 && ps.max_len >= t.d + 15 && !std::getenv("MK_NO_LONG_SEEDS");
    // stride 8, ASCII, L2-resident filter: mk_scan_dual8 takes its keys in the permuted window layout (the short key
    // is the window with the bases after the q-th masked away); the other L2-resident flavours use the ordered packing
    t.dual_perm = !want_smem && enc == 0 && t.d == 8 && t.q >= 12 && !std::getenv("MK_NO_DUAL8");
    if (t.dual_perm) window_masks(enc, t.q, &t.win_mask0, &t.win_mask1);
    if (win_layout && !want_smem) {
        win_layout = false;
        if (!long_seeds && !t.dual_perm) index_seeds(0);  // (with dual_perm the codes of the first pass are already right)
        TBT("index2");
    } else if (t.dual_perm && !long_seeds) {
        index_seeds(0);  // MK_NO_WIN_SCAN (a test switch): the first pass packed ordered codes, mk_scan_dual8 wants window codes
    }
    t.win = win_layout;

    // Seed sets too large for shared memory with a stride below 16: give every pattern that is long
    // enough a 16-base seed (far fewer text positions carry one of those than one of the q-base seeds)
    // (stride >= 2: the scan flags candidates in bit 0 of their position)
    if (long_seeds) {
        t.q2 = 16;
        t.long_min_len = t.d + 15;
        index_seeds(t.long_min_len);
    TBT("index3");
    }

    t.postings.resize(trips.size());
    for (size_t i = 0; i < trips.size(); ++i) {
        bool last = (i + 1 == trips.size()) || ((trips[i + 1] ^ trips[i]) >> 31) != 0;
        t.postings[i] = make_posting(t_pid(trips[i]), t_j(trips[i]), last);
    }
    if (t.postings.empty()) t.postings.push_back(make_posting(0, 0, true));  // keep device pointers non-null
    t.n_seeds = (uint32_t)keys.size();
    TBT("postings");

    // cuckoo table (4-slot buckets) at <= 80 % load, grown until the insertion succeeds
    uint32_t lb = 1;
    while (((uint64_t)kBucketSlots << lb) * 4 < (uint64_t)keys.size() * 5) ++lb;
    while (!cuckoo_build(keys, lb, &t.slots, &t.bucket_mask)) {
        if (++lb > 28) throw std::runtime_error("seed table does not fit");
    }

    TBT("cuckoo");
    const double nn = (double)t.n_seeds;
    // Strides 1 and 2 (shortest pattern below 15 bases), ASCII: a direct bitmap over the first q1 = min(q, 10) bases of
    // every seed instead of the hashed filter (mk_scan_short) — while it stays selective: at most 1 % of the 4^q1
    // prefixes occupied, or q <= 10 (then it is exact and a Bloom filter could only pass more). The second-level bitmap
    // on the full seed code is built as usual.
    if (want_smem && enc == 0 && t.d <= 2 && !t.win && !std::getenv("MK_NO_DIRECT")) {
        const uint32_t q1 = std::min<uint32_t>(t.q, 10);
        const uint32_t words = std::max<uint32_t>(4u, (uint32_t)(((uint64_t)1 << (2 * q1)) / 32));
        std::vector<uint32_t> bm(words, 0);
        size_t occupied = 0;
        for (auto& kv : keys) {
            const uint32_t idx = kv.code >> (2 * (t.q - q1));
            uint32_t& w = bm[idx >> 5];
            if (!((w >> (idx & 31)) & 1u)) { w |= 1u << (idx & 31); ++occupied; }
        }
        if (t.q <= 10 || (double)occupied <= 0.01 * std::ldexp(1.0, 2 * (int)q1)) {
            t.filter_direct = true;
            t.direct_q1 = q1;
        }
        if (t.filter_direct) {
            t.filter_in_smem = true;
            t.filter32 = false;
            t.filter_blocks = 0;
            t.filter_log2_bits = 0;
            t.filter_hashes = 1;
            t.filter = std::move(bm);
            uint32_t b2 = 16;
            while (b2 < 30 && std::ldexp(1.0, b2) < nn * 64.0) ++b2;
            t.filter2_log2_bits = b2;
            t.filter2.assign((size_t)1 << (b2 - 5), 0);
            for (auto& kv : keys) {
                uint32_t h = mk_hash_f2(kv.code, b2);
                t.filter2[h >> 5] |= 1u << (h & 31);
            }
        }
    }
    if (t.filter_direct) {
        // (tables done above)
    } else if (want_smem) {
        t.filter_in_smem = true;
        // small seed sets: 32-bit blocks are as selective and cost half the shared-memory traffic per probe
        // (k = 19..30, 2 000 patterns: 6.4 -> 7.0 TB/s); larger sets need the 64-bit blocks' lower false-positive rate
        double f32_max_fp = 0.004;
        if (const char* mf = std::getenv("MK_F32_MAX_FP")) f32_max_fp = std::atof(mf);  // tuning override
        t.filter32 = (t.perm || t.win) && blocked_fp32((double)keys.size(), 2 * nblocks) <= f32_max_fp && !std::getenv("MK_NO_FILTER32");
#ifdef MK_TUNE_BUILD
        t.filter32 = false;  // the launch-shape sweep only builds the 64-bit flavour
#endif
        t.filter_blocks = nblocks;
        t.filter_log2_bits = 0;
        t.filter_hashes = 4;
        t.filter.assign((size_t)nblocks * 2, 0);
        for (auto& kv : keys) {
            if (t.filter32) {
                t.filter[mk_bloom32_block(kv.code, nblocks)] |= mk_bloom32_mask(kv.code);
                continue;
            }
            uint32_t blk = mk_bloom_block(kv.code, nblocks), lo, hi;
            mk_bloom_masks(kv.code, &lo, &hi);
            t.filter[2 * (size_t)blk] |= lo;
            t.filter[2 * (size_t)blk + 1] |= hi;
        }
        uint32_t b2 = 16;
        while (b2 < 30 && std::ldexp(1.0, b2) < nn * 64.0) ++b2;
        t.filter2_log2_bits = b2;
        t.filter2.assign((size_t)1 << (b2 - 5), 0);
        for (auto& kv : keys) {
            uint32_t h = mk_hash_f2(kv.code, b2);
            t.filter2[h >> 5] |= 1u << (h & 31);
        }
    } else if (t.d < 16) {
        // L2-resident blocked Bloom, ~32 bits per key, 4 bits per key inside one 64-bit block that the
        // short and the long key of a text position share
        t.filter_in_smem = false;
        t.filter_dual = true;
        t.filter_hashes = 4;
        // bits per key: 32 (4-bit entries); 48 for mk_scan_dual8's 3-bit entries (cfg5: 3.1 M candidates at 32, 2.2 M at 48,
        // same scan time, verification 0.27 -> 0.21 ms)
        uint64_t nb = std::max<uint64_t>(1u << 15, t.dual_perm ? (uint64_t)t.n_seeds * 3 / 4 : (uint64_t)t.n_seeds / 2);
        if (const char* bk = std::getenv("MK_DUAL_BITS_PER_KEY")) nb = std::max<uint64_t>(1u << 15, (uint64_t)t.n_seeds * (uint64_t)std::atoi(bk) / 64);
        t.filter_blocks = (uint32_t)std::min<uint64_t>(nb, 1u << 27) & ~1u;
        t.filter.assign((size_t)t.filter_blocks * 2, 0);
        const uint32_t sshift = 32u - 2u * t.q;
        for (size_t ki = 0; ki < keys.size(); ++ki) {
            if (ki + 24 < keys.size()) {  // the filter is far larger than the cache
                const SeedKey& nx = keys[ki + 24];
                __builtin_prefetch(&t.filter[2 * (size_t)mk_dual_block(nx.group ? (t.dual_perm ? (nx.code & t.win_mask0) : (nx.code >> sshift)) : nx.code, t.filter_blocks)], 1);
            }
            const SeedKey& kv = keys[ki];
            uint32_t short_code = kv.group ? (t.dual_perm ? (kv.code & t.win_mask0) : (kv.code >> sshift)) : kv.code;
            uint32_t blk = mk_dual_block(short_code, t.filter_blocks), lo, hi;
            if (t.dual_perm) mk_dual3_masks(kv.group ? mk_dual3_g_long(kv.code) : mk_dual3_g_short(short_code), &lo, &hi);
            else mk_bloom_masks_g(kv.group ? mk_dual_g_long(kv.code) : mk_dual_g_short(short_code), &lo, &hi);
            t.filter[2 * (size_t)blk] |= lo;
            t.filter[2 * (size_t)blk + 1] |= hi;
        }
    } else {
        // stride 16, too many seeds for shared memory: L2-resident bitmap, ~32 bits per seed, one hash
        t.filter_in_smem = false;
        t.filter_hashes = 1;
        uint32_t b = 21;
        while (b < 30 && std::ldexp(1.0, b) < nn * 32.0) ++b;
        t.filter_log2_bits = b;
        t.filter.assign((size_t)1 << (t.filter_log2_bits - 5), 0);
        for (auto& kv : keys) {
            uint32_t h = mk_hash_f1(kv.code, t.filter_log2_bits);
            t.filter[h >> 5] |= 1u << (h & 31);
        }
    }
    TBT("filter");
    if (enc == 0 && !std::getenv("MK_NO_GATE")) {
        // bits that are the same in every byte a pattern can match
        bool in_set[256] = {false};
        for (size_t i = 0; i < ps.bytes.size(); ++i) in_set[t.pat_bytes[i]] = true;  // compare form: folded under -I
        if (ps.case_insensitive)
            for (int b = 'a'; b <= 'z'; ++b) if (in_set[b]) in_set[b - 0x20] = true;
        uint32_t all_and = 0xFF, all_or = 0;
        for (int b = 0; b < 256; ++b) if (in_set[b]) { all_and &= (uint32_t)b; all_or |= (uint32_t)b; }
        const uint32_t constant = (all_and | ~all_or) & 0xFF;  // bits that are 1 everywhere or 0 everywhere
        t.gate_mask = constant * 0x01010101u;
        t.gate_val = (all_and & constant) * 0x01010101u;
    }
    return t;
}

}  // namespace mk
