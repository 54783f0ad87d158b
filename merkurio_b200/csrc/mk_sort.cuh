// Hit-list kernels: stable LSD radix sort into the reference's report order, and the conversion of
// raw hits into mk_hit (de-duplicated for MK_MODE_PATTERN_SET).
//
// The list length lives on the device (min(*hit_count, capacity)); every kernel derives the same
// chunking from it, so the whole pipeline is enqueued without a host round trip. Lists are small
// (about one hit per hundred reads), so the kernels are tuned for launch count and latency:
// 8 bits per pass, only as many passes as the key has bits, three launches per pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/merkurio_cuda.h"
#include "mk_scan.cuh"

#ifndef MK_ASSERT
#ifdef MK_DEBUG_CHECKS
#include <cassert>
#define MK_ASSERT(c) assert(c)
#else
#define MK_ASSERT(c) ((void)0)
#endif
#endif

namespace mk {

constexpr int kSortBlocks = 256;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortBlocks * (kSortThreads / 32);  // 2048 chunk owners at most
constexpr uint32_t kSortChunkMin = 512;

struct SortGeom {
    uint64_t n;
    uint32_t g;      // active chunks (one warp each)
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

// pass 1/3: per-chunk digit histogram -> table[digit][chunk]
__global__ void __launch_bounds__(kSortThreads) mk_radix_hist(const RawHit* __restrict__ in, const unsigned long long* count,
                                                             unsigned long long cap, uint32_t shift, uint32_t* __restrict__ table) {
    __shared__ uint32_t cnt[kSortThreads / 32][256];
    const SortGeom sg = sort_geom(count, cap);
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (kSortThreads / 32) + w;
    if (wid >= sg.g) return;
    for (int i = lane; i < 256; i += 32) cnt[w][i] = 0;
    __syncwarp();
    uint64_t b = (uint64_t)wid * sg.chunk, e = b + sg.chunk;
    if (e > sg.n) e = sg.n;
    for (uint64_t i = b + lane; i < e; i += 32) atomicAdd(&cnt[w][(in[i].key >> shift) & 255u], 1u);
    __syncwarp();
    for (int dgt = lane; dgt < 256; dgt += 32) table[(size_t)dgt * sg.g + wid] = cnt[w][dgt];
}

// pass 2/3: one block per digit: exclusive scan of the digit's row over the chunks (in place) and
// the digit total into totals[digit]
__global__ void __launch_bounds__(256) mk_radix_rowscan(const unsigned long long* count, unsigned long long cap,
                                                       uint32_t* __restrict__ table, uint32_t* __restrict__ totals) {
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry;
    const SortGeom sg = sort_geom(count, cap);
    uint32_t* row = table + (size_t)blockIdx.x * sg.g;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < sg.g; base += 256) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < sg.g ? row[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t pre = carry;
        for (uint32_t k = 0; k < w; ++k) pre += warp_sum[k];
        if (i < sg.g) row[i] = pre + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = pre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// pass 3/3: each chunk owner scans the 256 digit totals itself, adds its row offsets and scatters
// its elements in order (stable)
__global__ void __launch_bounds__(kSortThreads) mk_radix_scatter(const RawHit* __restrict__ in, RawHit* __restrict__ out,
                                                                const unsigned long long* count, unsigned long long cap,
                                                                uint32_t shift, const uint32_t* __restrict__ table,
                                                                const uint32_t* __restrict__ totals) {
    __shared__ uint32_t base[kSortThreads / 32][256];
    const SortGeom sg = sort_geom(count, cap);
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (kSortThreads / 32) + w;
    if (wid >= sg.g) return;
    {   // exclusive scan of the digit totals: lane l owns digits 8l .. 8l+7
        uint32_t t[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t[k] = totals[lane * 8 + k]; s += t[k]; }
        uint32_t x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        uint32_t run = x - s;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            base[w][lane * 8 + k] = run + table[(size_t)(lane * 8 + k) * sg.g + wid];
            run += t[k];
        }
    }
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

// ---------------------------------------------------------------------------------------------
// Bucket sort: the usual hit list is a few hundred thousand entries whose keys are spread evenly
// (hits are scattered over the batch), so one MSD split into kBuckets ranges followed by a sort of
// each range in shared memory replaces five radix passes (15 launches) by 4 launches. A range with
// more than kBucketMax entries (hits piled up in one region) raises a flag; the host then runs the
// radix sort on the untouched input instead, and keeps doing so for that workspace.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kBucketBits = 13;
constexpr uint32_t kBuckets = 1u << kBucketBits;
constexpr uint32_t kBucketMax = 2048;

// Bucket of a key: monotonic in the key and spread over the keys that can actually occur (the largest
// key is rarely close to a power of two): the top 32 bits of the key range, scaled by `mult` / 2^32.
struct BucketMap {
    uint32_t shift;  // key >> shift fits 32 bits
    uint32_t mult;   // bucket = ((key >> shift) * mult) >> 32
};
__host__ __device__ __forceinline__ uint32_t bucket_of(unsigned long long key, BucketMap bm) {
    uint32_t b = (uint32_t)((((unsigned long long)(uint32_t)(key >> bm.shift)) * bm.mult) >> 32);
    return b < kBuckets ? b : kBuckets - 1;
}
inline BucketMap make_bucket_map(unsigned long long max_key) {
    BucketMap bm;
    bm.shift = 0;
    while ((max_key >> bm.shift) >> 32) ++bm.shift;
    const unsigned long long top = (max_key >> bm.shift) + 1;  // number of distinct key tops, <= 2^32
    unsigned long long mult = (((unsigned long long)kBuckets) << 32) / top;
    if (mult > 0xFFFFFFFFull) mult = 0xFFFFFFFFull;  // tiny key ranges: at most one bucket per key top (mult saturates)
    bm.mult = (uint32_t)mult;
    return bm;
}

__global__ void __launch_bounds__(256) mk_bucket_count(const RawHit* __restrict__ in, const unsigned long long* count,
                                                      unsigned long long cap, BucketMap bm, uint32_t* __restrict__ bucket_count) {
    unsigned long long n = *count;
    if (n > cap) n = cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        MK_ASSERT(bucket_of(in[i].key, bm) < (uint32_t)kBuckets);
        atomicAdd(&bucket_count[bucket_of(in[i].key, bm)], 1u);
    }
}

// one block: bucket_start = exclusive scan of bucket_count (kBuckets + 1 entries), cursor = copy of it,
// bucket_count zeroed for the next batch, *overflow = 1 if a bucket exceeds kBucketMax
__global__ void __launch_bounds__(1024) mk_bucket_scan(uint32_t* __restrict__ bucket_count, uint32_t* __restrict__ bucket_start,
                                                      uint32_t* __restrict__ cursor, unsigned long long* overflow) {
    __shared__ uint32_t warp_sum[32];
    constexpr int PER = kBuckets / 1024;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t v[PER], s = 0, big = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        v[k] = bucket_count[threadIdx.x * PER + k];
        bucket_count[threadIdx.x * PER + k] = 0;
        s += v[k];
        big |= v[k] > kBucketMax;
    }
    uint32_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[w] = x;
    const int any_big = __syncthreads_or((int)big);
    uint32_t pre = 0;
    for (uint32_t k = 0; k < w; ++k) pre += warp_sum[k];
    uint32_t run = pre + x - s;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        bucket_start[threadIdx.x * PER + k] = run;
        cursor[threadIdx.x * PER + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 1023) bucket_start[kBuckets] = run;
    if (threadIdx.x == 0) *overflow = any_big ? 1ull : 0ull;
}

__global__ void __launch_bounds__(256) mk_bucket_scatter(const RawHit* __restrict__ in, RawHit* __restrict__ out,
                                                        const unsigned long long* count, unsigned long long cap, BucketMap bm,
                                                        uint32_t* __restrict__ cursor, const unsigned long long* overflow) {
    if (*overflow) return;
    unsigned long long n = *count;
    if (n > cap) n = cap;
    // two hits per thread and step: the cursor atomics return a value (a full L2 round trip each), so the
    // kernel lives on how many of them are in flight
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
        const uint64_t j = i + stride;
        RawHit a = in[i], b;
        if (j < n) b = in[j];
        const uint32_t pa = atomicAdd(&cursor[bucket_of(a.key, bm)], 1u);
        uint32_t pb = 0;
        if (j < n) pb = atomicAdd(&cursor[bucket_of(b.key, bm)], 1u);
        MK_ASSERT(pa < n && (j >= n || pb < n));
        out[pa] = a;
        if (j < n) out[pb] = b;
    }
}

// Bitonic sort of the bucket by key in shared memory, then either the conversion to mk_hit (ALL_HITS:
// hits != nullptr) or the sorted raw hits written back in place (PATTERN_SET, de-duplicated afterwards by
// mk_heads_* / mk_finalize_pairs). Buckets of up to kWarpBucket entries — nearly all — are sorted by one
// warp each (8 at a time per block, only __syncwarp between the stages); the few larger ones by whole blocks.
constexpr uint32_t kWarpBucket = 256;
static_assert(kBucketMax == 8 * kWarpBucket, "the block path reuses the eight warp regions as one");

template <bool WARP>
__device__ __forceinline__ void bucket_sort_one(RawHit* __restrict__ data, uint32_t lo, uint32_t m, unsigned long long* s_key, uint2* s_pay,
                                                mk_hit* __restrict__ hits, const unsigned long long* __restrict__ off,
                                                const uint32_t* __restrict__ pat_off, uint32_t key_shift) {
    const uint32_t tid = WARP ? (threadIdx.x & 31) : threadIdx.x, nt = WARP ? 32 : blockDim.x;
    auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    uint32_t p2 = 1;
    while (p2 < m) p2 <<= 1;
    for (uint32_t i = tid; i < p2; i += nt) {
        if (i < m) {
            RawHit h = data[lo + i];
            s_key[i] = h.key;
            s_pay[i] = make_uint2(h.record, h.pattern);
        } else {
            s_key[i] = ~0ull;
        }
    }
    sync();
    for (uint32_t k = 2; k <= p2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            // one compare-exchange per thread and step: pair t is (i, i | j) with bit j of i clear
            for (uint32_t t = tid; t < (p2 >> 1); t += nt) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long a = s_key[i], c = s_key[l];
                if ((a > c) == up) {
                    s_key[i] = c; s_key[l] = a;
                    const uint2 x = s_pay[i]; s_pay[i] = s_pay[l]; s_pay[l] = x;
                }
            }
            sync();
        }
    }
    for (uint32_t i = tid; i < m; i += nt) {
        const uint2 pr = s_pay[i];
        if (hits) {
            const uint32_t L = pat_off[pr.y + 1] - pat_off[pr.y];
            mk_hit o;
            o.record = pr.x;
            o.start = (uint32_t)((s_key[i] >> key_shift) - L - off[pr.x]);
            o.pattern = pr.y;
            o.len = L;
            hits[lo + i] = o;
        } else {
            RawHit h;
            h.key = s_key[i];
            h.record = pr.x;
            h.pattern = pr.y;
            data[lo + i] = h;
        }
    }
    sync();
}

__global__ void __launch_bounds__(256) mk_bucket_sort(RawHit* __restrict__ data, const uint32_t* __restrict__ bucket_start,
                                                     const unsigned long long* overflow, mk_hit* __restrict__ hits,
                                                     const unsigned long long* __restrict__ off, const uint32_t* __restrict__ pat_off,
                                                     uint32_t key_shift) {
    __shared__ unsigned long long s_key[kBucketMax];
    __shared__ uint2 s_pay[kBucketMax];
    if (*overflow) return;
    const uint32_t w = threadIdx.x >> 5, warps = blockDim.x >> 5;
    // small buckets: one warp each
    for (uint32_t b = blockIdx.x * warps + w; b < kBuckets; b += gridDim.x * warps) {
        const uint32_t lo = bucket_start[b], m = bucket_start[b + 1] - lo;
        if (m == 0 || m > kWarpBucket) continue;
        bucket_sort_one<true>(data, lo, m, s_key + w * kWarpBucket, s_pay + w * kWarpBucket, hits, off, pat_off, key_shift);
    }
    __syncthreads();
    // large buckets: one block each
    for (uint32_t b = blockIdx.x; b < kBuckets; b += gridDim.x) {
        const uint32_t lo = bucket_start[b], m = bucket_start[b + 1] - lo;
        if (m <= kWarpBucket) continue;
        bucket_sort_one<false>(data, lo, m, s_key, s_pay, hits, off, pat_off, key_shift);
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

// ---------------------------------------------------------------------------------------------
// PATTERN_SET: keep the first hit of every (record, pattern) run of the sorted list.
// Tiles of 2048 elements; count heads per tile -> scan the tile counts (one block) -> every tile
// ranks its heads with a block scan and writes them.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kHeadTile = 2048;  // 256 threads x 8 consecutive elements

__device__ __forceinline__ uint32_t head_flags8(const RawHit* __restrict__ in, uint64_t first, uint64_t n, uint32_t* flags) {
    uint32_t c = 0;
    unsigned long long prev = first > 0 && first <= n ? in[first - 1].key : 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint64_t i = first + k;
        uint32_t f = 0;
        if (i < n) {
            unsigned long long key = in[i].key;
            f = (i == 0 || key != prev) ? 1u : 0u;
            prev = key;
        }
        flags[k] = f;
        c += f;
    }
    return c;
}

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total) {
    __shared__ uint32_t ws[8];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // protect ws against the previous use
    if (lane == 31) ws[w] = x;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        if (k < w) pre += ws[k];
        tot += ws[k];
    }
    *total = tot;
    return pre + x - v;
}

__global__ void __launch_bounds__(256) mk_heads_count(const RawHit* __restrict__ in, const unsigned long long* count,
                                                     unsigned long long cap, uint32_t* __restrict__ tile_count) {
    unsigned long long n = *count;
    if (n > cap) n = cap;
    const uint64_t tiles = (n + kHeadTile - 1) / kHeadTile;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint32_t f[8], total;
        uint32_t c = head_flags8(in, t * kHeadTile + threadIdx.x * 8ull, n, f);
        block_exclusive_scan_256(c, &total);
        if (threadIdx.x == 0) tile_count[t] = total;
    }
}

// exclusive scan of the tile counts (one block); total number of distinct pairs -> *n_out
__global__ void __launch_bounds__(1024) mk_heads_scan(const unsigned long long* count, unsigned long long cap,
                                                     uint32_t* __restrict__ tile_count, unsigned long long* n_out) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    unsigned long long n = *count;
    if (n > cap) n = cap;
    const uint64_t tiles = (n + kHeadTile - 1) / kHeadTile;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < tiles; base += 1024) {
        uint64_t i = base + threadIdx.x;
        uint32_t v = i < tiles ? tile_count[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t pre = carry;
        for (uint32_t k = 0; k < w; ++k) pre += warp_sum[k];
        if (i < tiles) tile_count[i] = pre + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = pre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}

__global__ void __launch_bounds__(256) mk_finalize_pairs(const RawHit* __restrict__ in, mk_hit* __restrict__ out,
                                                        const unsigned long long* count, unsigned long long cap,
                                                        const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ pat_off) {
    unsigned long long n = *count;
    if (n > cap) n = cap;
    const uint64_t tiles = (n + kHeadTile - 1) / kHeadTile;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint32_t f[8], total;
        const uint64_t first = t * kHeadTile + threadIdx.x * 8ull;
        uint32_t c = head_flags8(in, first, n, f);
        uint32_t dst = tile_off[t] + block_exclusive_scan_256(c, &total);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (f[k]) {
                RawHit h = in[first + k];
                mk_hit o;
                o.record = h.record;
                o.start = 0;
                o.pattern = h.pattern;
                o.len = pat_off[h.pattern + 1] - pat_off[h.pattern];
                out[dst++] = o;
            }
        }
    }
}

}  // namespace mk
