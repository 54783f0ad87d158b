// Hit-list kernels: stable LSD radix sort into the reference's report order, and the conversion of
// raw hits into mk_hit (de-duplicated for MK_MODE_PATTERN_SET).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/merkurio_cuda.h"
#include "mk_scan.cuh"

namespace mk {

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
