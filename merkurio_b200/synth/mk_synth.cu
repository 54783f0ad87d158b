// Synthetic read generator (see mk_synth.h). One implementation, compiled for host and device.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "mk_synth.h"

#define HD __host__ __device__ __forceinline__

namespace {

thread_local std::string g_err;

HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

HD uint32_t words_per_read(const mks_params& p) { return (p.read_len + 31) / 32; }

HD uint32_t raw_base(const mks_params& p, uint64_t r, uint32_t i) {
    uint64_t h = splitmix64(p.seed + r * words_per_read(p) + i / 32);
    return (uint32_t)(h >> (2 * (i % 32))) & 3u;
}

HD uint8_t base_char(uint32_t b) { return (uint8_t)("ACGT"[b]); }
HD uint8_t comp_char(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }

struct ReadPlan {
    bool plant, rc, nrun;
    uint32_t q, ppos, npos, nlen;
};

HD ReadPlan plan_read(const mks_params& p, uint64_t r) {
    ReadPlan pl;
    uint64_t h = splitmix64(p.seed ^ 0xD1B54A32D192ED03ull ^ (r * 0x2545F4914F6CDD1Dull));
    uint64_t h2 = splitmix64(h);
    pl.plant = p.n_queries && p.read_len >= p.k && (uint32_t)(h & 0xFFFF) < p.plant_per_65536;
    pl.q = p.n_queries ? (uint32_t)((h >> 16) % p.n_queries) : 0;
    pl.rc = (h >> 40) & 1;
    pl.ppos = p.read_len >= p.k ? (uint32_t)((h >> 41) % (p.read_len - p.k + 1)) : 0;
    pl.nrun = (uint32_t)(h2 & 0xFFFF) < p.nrun_per_65536 && p.read_len >= 10;
    pl.nlen = 1 + (uint32_t)((h2 >> 16) % 10);
    pl.npos = p.read_len >= 10 ? (uint32_t)((h2 >> 32) % (p.read_len - pl.nlen + 1)) : 0;
    return pl;
}

HD uint8_t read_char(const mks_params& p, const uint8_t* queries, uint64_t r, uint32_t i, const ReadPlan& pl) {
    if (pl.nrun && i >= pl.npos && i < pl.npos + pl.nlen) return 'N';
    if (pl.plant && i >= pl.ppos && i < pl.ppos + p.k) {
        uint32_t j = i - pl.ppos;
        const uint8_t* q = queries + (size_t)pl.q * p.k;
        return pl.rc ? comp_char(q[p.k - 1 - j]) : q[j];
    }
    return base_char(raw_base(p, r, i));
}

HD uint8_t nibble_of(uint8_t c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15; }

// one thread = 16 output bytes
__global__ void fill_kernel(mks_params p, const uint8_t* __restrict__ queries, uint64_t r0, uint64_t n_out_bytes, int enc,
                            uint4* __restrict__ out) {
    const uint32_t upb = enc ? 2 : 1;  // units per byte
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v * 16 < n_out_bytes; v += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t g = v * 16 * upb;  // first unit
        uint64_t r = r0 + g / p.read_len;
        uint32_t i = (uint32_t)(g % p.read_len);
        ReadPlan pl = plan_read(p, r);
        uint32_t w[4] = {0, 0, 0, 0};
        for (uint32_t b = 0; b < 16; ++b) {
            uint32_t byte = 0;
            if (v * 16 + b < n_out_bytes) {
                for (uint32_t u = 0; u < upb; ++u) {
                    uint8_t c = read_char(p, queries, r, i, pl);
                    byte = enc ? ((byte << 4) | nibble_of(c)) : c;
                    if (++i == p.read_len) { i = 0; ++r; pl = plan_read(p, r); }
                }
            }
            w[b >> 2] |= byte << (8 * (b & 3));
        }
        out[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void off_kernel(uint64_t n, uint32_t read_len, unsigned long long* off) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x)
        off[i] = i * read_len;
}

int check(const mks_params* p, int enc) {
    if (!p || p->read_len == 0) { g_err = "bad parameters"; return -1; }
    if (enc && (p->read_len & 1)) { g_err = "BAM4 output needs an even read length"; return -1; }
    if (p->n_queries && p->k > p->read_len) { g_err = "k > read_len"; return -1; }
    return 0;
}

}  // namespace

extern "C" {

const char* mks_last_error(void) { return g_err.c_str(); }

int mks_queries(const mks_params* p, uint8_t* out) {
    if (check(p, 0)) return -1;
    for (uint32_t j = 0; j < p->n_queries; ++j) {
        uint64_t r = ((uint64_t)j * 99991ull) % p->n_reads;
        uint32_t pos = j % (p->read_len - p->k + 1);
        for (uint32_t i = 0; i < p->k; ++i) out[(size_t)j * p->k + i] = base_char(raw_base(*p, r, pos + i));
    }
    return 0;
}

int mks_fill_host(const mks_params* p, const uint8_t* queries, uint64_t r0, uint64_t r1, int enc, uint8_t* out, uint64_t* off) {
    if (check(p, enc)) return -1;
    size_t w = 0;
    for (uint64_t r = r0; r < r1; ++r) {
        ReadPlan pl = plan_read(*p, r);
        if (off) off[r - r0] = (r - r0) * p->read_len;
        if (!enc) {
            for (uint32_t i = 0; i < p->read_len; ++i) out[w++] = read_char(*p, queries, r, i, pl);
        } else {
            for (uint32_t i = 0; i < p->read_len; i += 2)
                out[w++] = (uint8_t)((nibble_of(read_char(*p, queries, r, i, pl)) << 4) | nibble_of(read_char(*p, queries, r, i + 1, pl)));
        }
    }
    if (off) off[r1 - r0] = (r1 - r0) * p->read_len;
    return 0;
}

int mks_fill_device(const mks_params* p, const uint8_t* d_queries, uint64_t r0, uint64_t r1, int enc, void* d_out, uint64_t* d_off,
                    void* stream) {
    if (check(p, enc)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t n = r1 - r0;
    uint64_t bytes = enc ? n * p->read_len / 2 : n * p->read_len;
    if (bytes) {
        uint64_t vecs = (bytes + 15) / 16;
        int grid = (int)((vecs + 255) / 256 < 148 * 16 ? (vecs + 255) / 256 : 148 * 16);
        fill_kernel<<<grid, 256, 0, st>>>(*p, d_queries, r0, (bytes + 15) / 16 * 16, enc, (uint4*)d_out);
    }
    if (d_off) off_kernel<<<148 * 4, 256, 0, st>>>(n, p->read_len, (unsigned long long*)d_off);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -2; }
    return 0;
}

}  // extern "C"
