// Device side of the matching engine (sm_100a).
//
//   mk_scan_*      : the scan that replaces BNDMq::find_iter / find_match
//                    (/root/reference/src/pattern_matching.rs:128-209) and
//                    AhoCorasick::find_overlapping_iter (src/cmd_extract.rs:332,480,507,
//                    src/cmd_tag.rs:393-396). HBM-bound: every text byte is read once with coalesced
//                    16-byte loads, packed to 2-bit classes in registers, and one seed per D bases is
//                    probed in a bitmap staged in shared memory; the rare survivors go through the
//                    L2-resident cuckoo seed table and an exact byte compare.
//   mk_radix_*     : LSD radix sort of the compacted hit list into the reference's report order.
//   mk_finalize_*  : raw hits -> mk_hit, de-duplication for MK_MODE_PATTERN_SET.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/merkurio_cuda.h"
#include "mk_codes.h"
#include "mk_tables.h"

namespace mk {

struct RawHit {
    unsigned long long key;
    uint32_t record;
    uint32_t pattern;
};
static_assert(sizeof(RawHit) == 16, "RawHit must be 16 bytes");

struct ScanParams {
    // text
    const uint4* text;   // raw sequence bytes, 16-byte aligned
    uint64_t n_units;    // bases in the batch
    uint64_t n_vec;      // 16-byte vectors that cover them
    const unsigned long long* off;  // n_records + 1 unit offsets
    const uint32_t* lens;           // optional record lengths
    uint32_t n_records;
    // tables
    const uint32_t* filter;
    uint32_t filter_log2_bits;
    const SeedSlot* slots;
    uint32_t bucket_mask;
    const uint32_t* postings;
    const uint8_t* pat_bytes;
    const uint32_t* pat_off;
    const uint32_t* tie_rank;
    uint32_t q;
    int case_insensitive;
    // outputs
    uint32_t* flags;                // 1 bit / record
    RawHit* hits;
    unsigned long long hit_capacity;
    unsigned long long* hit_count;
    int mode;
    uint32_t len_bits, tie_bits, pat_bits, max_len;
};

constexpr int kScanThreads = 1024;
constexpr int kScanWarps = kScanThreads / 32;

// Streaming 16-byte load: read-only path, no L1 allocation, evict-first in L2 so that the seed
// tables and pattern bytes stay L2-resident while the batch streams through.
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

template <int NHASH>
__device__ __forceinline__ bool filter_probe(const uint32_t* __restrict__ f, uint32_t code, uint32_t lb) {
    uint32_t h = mk_hash_f1(code, lb);
    bool pass = (f[h >> 5] >> (h & 31)) & 1u;
    if (NHASH == 2) {
        uint32_t h2 = mk_hash_f2(code, lb);
        pass = pass && ((f[h2 >> 5] >> (h2 & 31)) & 1u);
    }
    return pass;
}

// last record r with off[r] <= s  (records of length 0 are skipped by construction)
__device__ __forceinline__ uint32_t find_record(const unsigned long long* __restrict__ off, uint32_t n, uint64_t s) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] <= s) lo = mid; else hi = mid;
    }
    return lo;
}

template <int ENC>
__device__ __forceinline__ uint8_t text_symbol(const uint8_t* __restrict__ t, uint64_t pos, int ci) {
    if (ENC == MK_ENC_ASCII) {
        uint8_t c = t[pos];
        return ci ? mk_fold(c) : c;
    }
    uint8_t b = t[pos >> 1];
    return (pos & 1) ? (b & 0xF) : (b >> 4);
}

// Slow path: a seed at base position `pos` passed the first-level filter.
template <int ENC>
__device__ __noinline__ void verify_seed(const ScanParams& P, uint64_t pos, uint32_t code) {
    // cuckoo lookup: two 32-byte buckets
    uint32_t first = kEmptySlot;
#pragma unroll
    for (int h = 0; h < 2 && first == kEmptySlot; ++h) {
        uint32_t b = h == 0 ? mk_hash_b1(code, P.bucket_mask) : mk_hash_b2(code, P.bucket_mask);
        const uint4* bp = reinterpret_cast<const uint4*>(P.slots + (size_t)b * kBucketSlots);
        uint4 lo = __ldg(bp), hi = __ldg(bp + 1);
        if (lo.x == code && lo.y != kEmptySlot) first = lo.y;
        else if (lo.z == code && lo.w != kEmptySlot) first = lo.w;
        else if (hi.x == code && hi.y != kEmptySlot) first = hi.y;
        else if (hi.z == code && hi.w != kEmptySlot) first = hi.w;
    }
    if (first == kEmptySlot) return;

    const uint8_t* text = reinterpret_cast<const uint8_t*>(P.text);
    for (uint32_t i = first;; ++i) {
        uint32_t e = __ldg(P.postings + i);
        uint32_t pid = e >> 5, j = (e >> 1) & 15u;
        if (pos >= j) {
            uint64_t s = pos - j;
            uint32_t po = __ldg(P.pat_off + pid);
            uint32_t L = __ldg(P.pat_off + pid + 1) - po;
            if (s + L <= P.n_units) {
                const uint8_t* pat = P.pat_bytes + po;
                bool eq = true;
                for (uint32_t k = 0; k < L; ++k) {
                    if (text_symbol<ENC>(text, s + k, P.case_insensitive) != __ldg(pat + k)) { eq = false; break; }
                }
                if (eq && s >= P.off[0]) {
                    uint32_t r = find_record(P.off, P.n_records, s);
                    uint64_t rend = P.lens ? P.off[r] + P.lens[r] : P.off[r + 1];
                    if (s + L <= rend) {
                        atomicOr(P.flags + (r >> 5), 1u << (r & 31));
                        if (P.mode != MK_MODE_FLAG) {
                            // warp-aggregated append: one atomic per group of lanes that got here together
                            cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
                            unsigned long long base = 0;
                            if (g.thread_rank() == 0) base = atomicAdd(P.hit_count, (unsigned long long)g.size());
                            base = g.shfl(base, 0);
                            unsigned long long slot = base + g.thread_rank();
                            if (slot < P.hit_capacity) {
                                RawHit hrec;
                                if (P.mode == MK_MODE_ALL_HITS)
                                    hrec.key = ((((unsigned long long)(s + L) << P.len_bits) | (P.max_len - L)) << P.tie_bits) |
                                               __ldg(P.tie_rank + pid);
                                else
                                    hrec.key = ((unsigned long long)r << P.pat_bits) | pid;
                                hrec.record = r;
                                hrec.pattern = pid;
                                P.hits[slot] = hrec;
                            }
                        }
                    }
                }
            }
        }
        if (e & 1u) break;
    }
}

template <bool SMEMF>
__device__ __forceinline__ const uint32_t* stage_filter(const ScanParams& P, uint32_t* s_filter) {
    if (!SMEMF) return P.filter;
    const uint32_t n16 = 1u << (P.filter_log2_bits - 7);  // uint4 count
    const uint4* src = reinterpret_cast<const uint4*>(P.filter);
    uint4* dst = reinterpret_cast<uint4*>(s_filter);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    return s_filter;
}

// ---------------------------------------------------------------------------------------------
// D == 16: the seed is the 16-base unit itself; no lane needs its neighbour.
// ASCII: one seed per 16-byte vector. BAM4: two seeds per vector.
// ---------------------------------------------------------------------------------------------
template <int ENC, int NHASH, bool SMEMF, int U>
__global__ void __launch_bounds__(kScanThreads, 1) mk_scan_d16(const __grid_constant__ ScanParams P) {
    extern __shared__ __align__(16) uint32_t s_filter[];
    const uint32_t* __restrict__ filt = stage_filter<SMEMF>(P, s_filter);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lb = P.filter_log2_bits;
    const uint64_t pol = make_evict_first_policy();
    const uint64_t nwarps = (uint64_t)gridDim.x * kScanWarps;
    const uint64_t n_vec = P.n_vec;
    constexpr int SPV = (ENC == MK_ENC_ASCII) ? 1 : 2;  // seeds (= units) per vector

    for (uint64_t tile = (uint64_t)blockIdx.x * kScanWarps + (threadIdx.x >> 5); tile * (U * 32) < n_vec; tile += nwarps) {
        const uint64_t v0 = tile * (U * 32) + lane;
        uint4 v[U];
        if ((tile + 1) * (U * 32) <= n_vec) {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_stream(P.text + v0 + u * 32, pol);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (v0 + u * 32 < n_vec) ? ld_stream(P.text + v0 + u * 32, pol) : make_uint4(0, 0, 0, 0);
        }
        uint32_t code[U * SPV];
        uint32_t pass = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ENC == MK_ENC_ASCII) {
                code[u] = mk_pack_ascii_perm(v[u].x, v[u].y, v[u].z, v[u].w);
                pass |= (uint32_t)filter_probe<NHASH>(filt, code[u], lb) << u;
            } else {
                code[2 * u] = mk_pack_bam_perm(v[u].x, v[u].y);
                code[2 * u + 1] = mk_pack_bam_perm(v[u].z, v[u].w);
                pass |= (uint32_t)filter_probe<NHASH>(filt, code[2 * u], lb) << (2 * u);
                pass |= (uint32_t)filter_probe<NHASH>(filt, code[2 * u + 1], lb) << (2 * u + 1);
            }
        }
        if (pass) {
#pragma unroll
            for (int k = 0; k < U * SPV; ++k) {
                if (pass & (1u << k)) {
                    uint64_t unit = (v0 + (uint64_t)(k / SPV) * 32) * SPV + (k % SPV);
                    if (unit * MK_UNIT_BASES < P.n_units) verify_seed<ENC>(P, unit * MK_UNIT_BASES, code[k]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// D < 16: ordered unit codes; a seed at offset o of a unit is a funnel shift of (unit, next unit).
// The next unit lives in the next lane (shuffle), in lane 0 of the next row, or in the first
// vector of the next tile (one extra load by lane 0).
// ---------------------------------------------------------------------------------------------
template <int ENC, int D, int NHASH, bool SMEMF, int U>
__global__ void __launch_bounds__(kScanThreads, 1) mk_scan_ord(const __grid_constant__ ScanParams P) {
    extern __shared__ __align__(16) uint32_t s_filter[];
    const uint32_t* __restrict__ filt = stage_filter<SMEMF>(P, s_filter);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lb = P.filter_log2_bits, q = P.q;
    const uint64_t pol = make_evict_first_policy();
    const uint64_t nwarps = (uint64_t)gridDim.x * kScanWarps;
    const uint64_t n_vec = P.n_vec;
    constexpr int SPV = (ENC == MK_ENC_ASCII) ? 1 : 2;  // units per vector

    for (uint64_t tile = (uint64_t)blockIdx.x * kScanWarps + (threadIdx.x >> 5); tile * (U * 32) < n_vec; tile += nwarps) {
        const uint64_t v0 = tile * (U * 32) + lane;
        uint32_t c[(U + 1) * SPV];  // ordered unit codes of this lane's vectors, + the halo vector (valid in lane 0)
#pragma unroll
        for (int u = 0; u <= U; ++u) {
            uint64_t idx = v0 + (uint64_t)u * 32;
            bool want = (u < U || lane == 0) && idx < n_vec;
            uint4 v = want ? ld_stream(P.text + idx, pol) : make_uint4(0, 0, 0, 0);
            if (ENC == MK_ENC_ASCII) {
                c[u] = mk_pack_ascii_ord(v.x, v.y, v.z, v.w);
            } else {
                c[2 * u] = mk_pack_bam_ord(v.x, v.y);
                c[2 * u + 1] = mk_pack_bam_ord(v.z, v.w);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // first unit code of the following vector
            uint32_t from_next_lane = __shfl_down_sync(0xFFFFFFFFu, c[u * SPV], 1);
            uint32_t from_next_row = __shfl_sync(0xFFFFFFFFu, c[(u + 1) * SPV], 0);
            uint32_t succ_vec = (lane == 31) ? from_next_row : from_next_lane;
#pragma unroll
            for (int h = 0; h < SPV; ++h) {
                uint32_t cur = c[u * SPV + h];
                uint32_t nxt = (h + 1 < SPV) ? c[u * SPV + h + 1] : succ_vec;
                uint64_t unit = (v0 + (uint64_t)u * 32) * SPV + h;
                uint64_t base = unit * MK_UNIT_BASES;
#pragma unroll 4
                for (int o = 0; o < MK_UNIT_BASES; o += D) {
                    uint32_t seed = mk_seed_ord(cur, nxt, o, q);
                    if (filter_probe<NHASH>(filt, seed, lb) && base + o < P.n_units) verify_seed<ENC>(P, base + o, seed);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Hit list: stable LSD radix sort, 8 bits per pass. The list length lives on the device
// (min(*hit_count, capacity)); every kernel derives the same chunking from it, so the whole
// pipeline is enqueued without a host round trip.
// ---------------------------------------------------------------------------------------------
constexpr int kSortBlocks = 128;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortBlocks * (kSortThreads / 32);  // 1024 chunk owners at most
constexpr uint32_t kSortChunkMin = 2048;

struct SortGeom {
    uint64_t n;
    uint32_t g;      // active chunks
    uint64_t chunk;  // elements per chunk (multiple of 32)
};
__device__ __forceinline__ SortGeom sort_geom(const unsigned long long* count, unsigned long long cap) {
    SortGeom s;
    unsigned long long n = *count;
    s.n = n < cap ? n : cap;
    uint64_t g = (s.n + kSortChunkMin - 1) / kSortChunkMin;
    if (g < 1) g = 1;
    if (g > kSortWarps) g = kSortWarps;
    s.g = (uint32_t)g;
    s.chunk = ((s.n + g - 1) / g + 31) & ~31ull;
    return s;
}

__global__ void __launch_bounds__(kSortThreads) mk_radix_hist(const RawHit* __restrict__ in, const unsigned long long* count,
                                                             unsigned long long cap, uint32_t shift, uint32_t* __restrict__ table) {
    __shared__ uint32_t cnt[kSortThreads / 32][256];
    const SortGeom sg = sort_geom(count, cap);
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (kSortThreads / 32) + w;
    for (int i = lane; i < 256; i += 32) cnt[w][i] = 0;
    __syncwarp();
    if (wid < sg.g) {
        uint64_t b = (uint64_t)wid * sg.chunk, e = b + sg.chunk;
        if (e > sg.n) e = sg.n;
        for (uint64_t i = b + lane; i < e; i += 32) atomicAdd(&cnt[w][(in[i].key >> shift) & 255u], 1u);
        __syncwarp();
        for (int dgt = lane; dgt < 256; dgt += 32) table[(size_t)dgt * sg.g + wid] = cnt[w][dgt];
    }
}

// exclusive scan of the digit-major table (256 * g entries), one block
__global__ void __launch_bounds__(1024) mk_radix_scan(const unsigned long long* count, unsigned long long cap, uint32_t* table) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const SortGeom sg = sort_geom(count, cap);
    const uint32_t total = 256u * sg.g;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < total; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < total ? table[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
                if (lane >= o) s += y;
            }
            warp_sum[lane] = s;  // inclusive
        }
        __syncthreads();
        uint32_t prefix = carry + (w ? warp_sum[w - 1] : 0) + (x - v);
        if (i < total) table[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads) mk_radix_scatter(const RawHit* __restrict__ in, RawHit* __restrict__ out,
                                                                const unsigned long long* count, unsigned long long cap,
                                                                uint32_t shift, const uint32_t* __restrict__ table) {
    __shared__ uint32_t base[kSortThreads / 32][256];
    const SortGeom sg = sort_geom(count, cap);
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (kSortThreads / 32) + w;
    if (wid >= sg.g) return;
    for (int dgt = lane; dgt < 256; dgt += 32) base[w][dgt] = table[(size_t)dgt * sg.g + wid];
    __syncwarp();
    uint64_t b = (uint64_t)wid * sg.chunk, e = b + sg.chunk;
    if (e > sg.n) e = sg.n;
    for (uint64_t i0 = b; i0 < e; i0 += 32) {
        uint64_t i = i0 + lane;
        bool live = i < e;
        RawHit h;
        uint32_t dgt = 0x10000u + lane;  // idle lanes: unique digits, never written
        if (live) { h = in[i]; dgt = (uint32_t)(h.key >> shift) & 255u; }
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, dgt);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t dst = 0;
        if (live) dst = base[w][dgt] + rank;
        __syncwarp();
        if (live) {
            out[dst] = h;
            if (rank == 0) base[w][dgt] += __popc(peers);
        }
        __syncwarp();
    }
}

// ALL_HITS: sorted raw hits -> mk_hit
__global__ void mk_finalize_hits(const RawHit* __restrict__ in, mk_hit* __restrict__ out, const unsigned long long* count,
                                 unsigned long long cap, const unsigned long long* __restrict__ off,
                                 const uint32_t* __restrict__ pat_off, uint32_t key_shift) {
    unsigned long long n = *count;
    if (n > cap) n = cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        RawHit h = in[i];
        uint32_t L = pat_off[h.pattern + 1] - pat_off[h.pattern];
        uint64_t gend = h.key >> key_shift;
        mk_hit o;
        o.record = h.record;
        o.start = (uint32_t)(gend - L - off[h.record]);
        o.pattern = h.pattern;
        o.len = L;
        out[i] = o;
    }
}

// PATTERN_SET: mark the first hit of every (record, pattern) run ...
__global__ void mk_mark_heads(const RawHit* __restrict__ in, const unsigned long long* count, unsigned long long cap,
                              uint32_t* __restrict__ head) {
    unsigned long long n = *count;
    if (n > cap) n = cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        head[i] = (i == 0 || in[i].key != in[i - 1].key) ? 1u : 0u;
}
// ... exclusive-scan the marks (one block), leaving the number of distinct pairs in *n_out ...
__global__ void __launch_bounds__(1024) mk_scan_heads(const unsigned long long* count, unsigned long long cap, uint32_t* head,
                                                      unsigned long long* n_out) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    unsigned long long n = *count;
    if (n > cap) n = cap;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        uint64_t i = base + threadIdx.x;
        uint32_t v = i < n ? head[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
                if (lane >= o) s += y;
            }
            warp_sum[lane] = s;
        }
        __syncthreads();
        uint32_t prefix = carry + (w ? warp_sum[w - 1] : 0) + (x - v);
        // keep the mark in the top bit, the exclusive rank below it
        if (i < n) head[i] = (v << 31) | prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}
// ... and write one mk_hit per distinct pair.
__global__ void mk_finalize_pairs(const RawHit* __restrict__ in, mk_hit* __restrict__ out, const unsigned long long* count,
                                  unsigned long long cap, const uint32_t* __restrict__ head,
                                  const uint32_t* __restrict__ pat_off) {
    unsigned long long n = *count;
    if (n > cap) n = cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t m = head[i];
        if (m >> 31) {
            RawHit h = in[i];
            mk_hit o;
            o.record = h.record;
            o.start = 0;
            o.pattern = h.pattern;
            o.len = pat_off[h.pattern + 1] - pat_off[h.pattern];
            out[m & 0x7FFFFFFFu] = o;
        }
    }
}

}  // namespace mk
