// Micro-benchmark behind DESIGN.md section 4 ("what bounds the cfg5 scan"): how many lane-divergent
// filter probes per cycle one SM sustains through each path a first-level filter could live behind:
//   lds      random 4-/8-byte loads from the CTA's own shared memory
//   dsmem    random 4-/8-byte loads from the shared memory of all CTAs of a thread-block cluster
//            (mapa + ld.shared::cluster), cluster sizes 2..16
//   l2       random 8-byte __ldg gathers from an L2-resident table (what mk_scan_ord does today)
//   tex      the same gathers through a texture object (tex1Dfetch)
//   mix      one dsmem load and one l2 gather per step (do the two paths overlap?)
//   tex+l2   half of the gathers through tex1Dfetch, half through __ldg (are the two front ends independent?)
//   l2x16B   16-byte gathers (does the width of a lane-divergent load matter?)
//   dst      random 8-byte remote stores (st.shared::cluster), the "send the probe to the owner" variant
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
// Output: one line per variant, probes per cycle and SM (cycles from clock64 inside the kernel).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

constexpr int kThreads = 1024;
constexpr int kUnroll = 8;

enum Mode { LDS4 = 0, LDS8, DSMEM4, DSMEM8, L2G, TEXG, MIX, DST8, TEXMIX, L2G16 };

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s ^ (s >> 15); }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t ld_cluster32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 ld_cluster64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared::cluster.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void st_cluster64(uint32_t a, uint2 v) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(v.x), "r"(v.y) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) bench(const uint2* __restrict__ table, uint32_t table_mask, cudaTextureObject_t tex,
                                                     uint32_t smem_words, uint32_t csize, int iters, unsigned long long* cycles,
                                                     uint32_t* sink) {
    extern __shared__ __align__(16) uint32_t sm[];
    for (uint32_t i = threadIdx.x; i < smem_words; i += kThreads) sm[i] = i * 2654435761u;
    cg::cluster_group cl = cg::this_cluster();
    if (csize > 1) cl.sync(); else __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 747796405u + 12345u;
    uint32_t acc = 0;
    const uint32_t wmask8 = smem_words / 2 - 1;  // smem_words is a power of two
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) r[u] = lcg(s);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (MODE == LDS4) acc ^= sm[r[u] & (smem_words - 1)];
            if (MODE == LDS8) { uint2 v = reinterpret_cast<uint2*>(sm)[r[u] & wmask8]; acc ^= v.x + v.y; }
            if (MODE == DSMEM4 || MODE == MIX) {
                uint32_t a = mapa(sbase + 4u * (r[u] & (smem_words - 1)), (r[u] >> 24) % csize);
                acc ^= ld_cluster32(a);
            }
            if (MODE == DSMEM8) {
                uint32_t a = mapa(sbase + 8u * (r[u] & wmask8), (r[u] >> 24) % csize);
                uint2 v = ld_cluster64(a);
                acc ^= v.x + v.y;
            }
            if (MODE == L2G16) { uint4 v = __ldg(reinterpret_cast<const uint4*>(table) + (((r[u] * 2246822519u) & table_mask) >> 1)); acc ^= v.x + v.y + v.z + v.w; }
            if (MODE == TEXMIX) {  // half of the probes through the texture path, half through LDG
                if (u & 1) { uint2 v = tex1Dfetch<uint2>(tex, (int)((r[u] * 2246822519u) & table_mask)); acc ^= v.x + v.y; }
                else { uint2 v = __ldg(table + ((r[u] * 2246822519u) & table_mask)); acc ^= v.x + v.y; }
            }
            if (MODE == L2G || MODE == MIX) { uint2 v = __ldg(table + ((r[u] * 2246822519u) & table_mask)); acc ^= v.x + v.y; }
            if (MODE == TEXG) { uint2 v = tex1Dfetch<uint2>(tex, (int)((r[u] * 2246822519u) & table_mask)); acc ^= v.x + v.y; }
            if (MODE == DST8) st_cluster64(mapa(sbase + 8u * (r[u] & wmask8), (r[u] >> 24) % csize), make_uint2(r[u], acc));
        }
    }
    const long long t1 = clock64();
    if (csize > 1) cl.sync(); else __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE>
void run(const char* name, int csize, int sms, const uint2* table, uint32_t table_mask, cudaTextureObject_t tex, uint32_t smem_words, int iters,
         unsigned long long* d_cycles, uint32_t* d_sink) {
    const size_t smem = (size_t)smem_words * 4;
    CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int grid = sms / csize * csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (csize > 1) {
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, bench<MODE>, &cfg);
        if (e != cudaSuccess || ncl == 0) { std::printf("%-8s cluster %2d: cannot launch (%s, %d clusters)\n", name, csize, cudaGetErrorString(e), ncl); cudaGetLastError(); return; }
        if (ncl * csize < grid) { grid = ncl * csize; cfg.gridDim = dim3(grid); }
    }
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        CK(cudaLaunchKernelEx(&cfg, bench<MODE>, table, table_mask, tex, smem_words, (uint32_t)csize, iters, d_cycles, d_sink));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    std::vector<unsigned long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), d_cycles, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double mean = 0; unsigned long long mx = 0;
    for (auto c : cyc) { mean += (double)c; if (c > mx) mx = c; }
    mean /= grid;
    const double probes_cta = (double)kThreads * iters * kUnroll * (MODE == MIX ? 2 : 1);
    std::printf("%-8s cluster %2d  grid %3d  smem %3zu KiB/CTA  %.3f ms  %.2f Gprobes/s  %.3f probes/cycle/SM (mean CTA), %.3f (slowest CTA)\n", name, csize, grid,
                smem / 1024, best, probes_cta * grid / best / 1e6, probes_cta / mean, probes_cta / (double)mx);
    std::fflush(stdout);
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? std::atoi(argv[1]) : 400;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint32_t table_elems = 1u << 22;  // 4 M x 8 bytes = 32 MiB: the size of cfg5's dual-key filter
    uint2* table; unsigned long long* d_cycles; uint32_t* d_sink;
    CK(cudaMalloc(&table, (size_t)table_elems * 8));
    CK(cudaMemset(table, 0x5A, (size_t)table_elems * 8));
    CK(cudaMalloc(&d_cycles, 1024 * 8));
    CK(cudaMalloc(&d_sink, 64));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<uint2>();
    rd.res.linear.sizeInBytes = (size_t)table_elems * 8;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const uint32_t W = 32768;  // 128 KiB of shared memory per CTA (power of two)
    std::printf("SMs %d, %d threads/CTA, %d x %d probes per thread\n", sms, kThreads, iters, kUnroll);
    run<LDS4>("lds4", 1, sms, table, table_elems - 1, tex, W, iters, d_cycles, d_sink);
    run<LDS8>("lds8", 1, sms, table, table_elems - 1, tex, W, iters, d_cycles, d_sink);
    run<L2G>("l2", 1, sms, table, table_elems - 1, tex, W, iters / 4, d_cycles, d_sink);
    run<TEXG>("tex", 1, sms, table, table_elems - 1, tex, W, iters / 4, d_cycles, d_sink);
    run<TEXMIX>("tex+l2", 1, sms, table, table_elems - 1, tex, W, iters / 4, d_cycles, d_sink);
    run<L2G16>("l2x16B", 1, sms, table, table_elems - 1, tex, W, iters / 4, d_cycles, d_sink);
    for (int c : {2, 4, 8, 16}) {
        run<DSMEM4>("dsmem4", c, sms, table, table_elems - 1, tex, W, iters / 2, d_cycles, d_sink);
        run<DSMEM8>("dsmem8", c, sms, table, table_elems - 1, tex, W, iters / 2, d_cycles, d_sink);
        run<DST8>("dst8", c, sms, table, table_elems - 1, tex, W, iters / 2, d_cycles, d_sink);
        run<MIX>("mix", c, sms, table, table_elems - 1, tex, W, iters / 4, d_cycles, d_sink);
    }
    return 0;
}
