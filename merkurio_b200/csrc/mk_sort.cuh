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
