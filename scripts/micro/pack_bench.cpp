// Micro-benchmark (host only): how fast can T threads pack ASCII bases to BAM 4-bit codes with AVX2?
// Decides whether packing on the host before the H2D copy can beat the PCIe rate of the ASCII stream.
//   g++ -O3 -pthread -o pack_bench pack_bench.cpp && ./pack_bench [MiB] [threads...]
#include <immintrin.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

// 32 ASCII bytes -> 16 packed bytes; returns a mask of bytes that are not in "=ACMGRSVTWYHKDBN"
__attribute__((target("avx2"))) static inline uint32_t pack32(const uint8_t* src, uint8_t* dst) {
    // nibble code by the low 5 bits of the character (unique inside the alphabet), 0xFF = not a member
    const __m256i lut_lo = _mm256_setr_epi8(-1, 1, 14, 2, 13, -1, -1, 4, 11, -1, -1, 12, -1, 3, 15, -1,
                                            -1, 1, 14, 2, 13, -1, -1, 4, 11, -1, -1, 12, -1, 3, 15, -1);
    const __m256i lut_hi = _mm256_setr_epi8(-1, -1, 5, 6, 8, -1, 7, 9, -1, 10, -1, -1, -1, 0, -1, -1,
                                            -1, -1, 5, 6, 8, -1, 7, 9, -1, 10, -1, -1, -1, 0, -1, -1);
    // the character each code stands for, to verify membership (bits 5..7 must match too)
    const __m256i chars = _mm256_setr_epi8('=', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N',
                                           '=', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N');
    __m256i c = _mm256_loadu_si256((const __m256i*)src);
    __m256i idx = _mm256_and_si256(c, _mm256_set1_epi8(0x0F));
    __m256i lo = _mm256_shuffle_epi8(lut_lo, idx), hi = _mm256_shuffle_epi8(lut_hi, idx);
    __m256i sel = _mm256_cmpeq_epi8(_mm256_and_si256(c, _mm256_set1_epi8(0x10)), _mm256_setzero_si256());
    __m256i code = _mm256_blendv_epi8(hi, lo, sel);
    __m256i back = _mm256_shuffle_epi8(chars, _mm256_and_si256(code, _mm256_set1_epi8(0x0F)));
    uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(back, c)) | (uint32_t)_mm256_movemask_epi8(code);
    // pairs (even byte -> high nibble): maddubs with (16, 1) then pack 16-bit lanes to bytes
    __m256i pairs = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0110));
    __m256i packed = _mm256_packus_epi16(pairs, pairs);                 // lanes: [p0..7 p0..7 | p8..15 p8..15]
    packed = _mm256_permute4x64_epi64(packed, 0x08);                     // gather qwords 0 and 2
    _mm_storeu_si128((__m128i*)dst, _mm256_castsi256_si128(packed));
    return bad;
}

int main(int argc, char** argv) {
    size_t mib = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1024;
    size_t n = mib << 20;
    std::vector<uint8_t> src(n), dst(n / 2 + 64);
    const char* al = "ACGTN";
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; src[i] = (uint8_t)al[x % 5]; }
    // correctness of the first block against the scalar rule
    {
        uint8_t out[16];
        uint32_t bad = pack32(src.data(), out);
        const char* codes = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; ++i) {
            int a = (int)(strchr(codes, src[2 * i]) - codes), b = (int)(strchr(codes, src[2 * i + 1]) - codes);
            if (out[i] != ((a << 4) | b)) { std::printf("MISMATCH at %d\n", i); return 1; }
        }
        uint8_t t[32];
        std::memcpy(t, src.data(), 32);
        t[5] = 'a';
        if (bad != 0 || !(pack32(t, out) & (1u << 5))) { std::printf("membership check wrong\n"); return 1; }
    }
    for (int a = 2; a < argc || a == 2; ++a) {
        int T = a < argc ? std::atoi(argv[a]) : (int)std::thread::hardware_concurrency();
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            std::vector<uint32_t> bad(T, 0);
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    size_t lo = n / T * t / 32 * 32, hi = t == T - 1 ? n : n / T * (t + 1) / 32 * 32;
                    uint32_t b = 0;
                    for (size_t i = lo; i + 32 <= hi; i += 32) b |= pack32(src.data() + i, dst.data() + i / 2);
                    bad[t] = b;
                });
            for (auto& t : th) t.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            best = std::max(best, (double)n / s / 1e9);
        }
        std::printf("%d threads: %.1f GB/s of ASCII packed\n", T, best);
        if (a >= argc) break;
    }
    return 0;
}
