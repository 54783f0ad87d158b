// Device scan of the matching engine (sm_100a): replaces BNDMq::find_iter / find_match
// (/root/reference/src/pattern_matching.rs:128-209) and AhoCorasick::find_overlapping_iter
// (src/cmd_extract.rs:332,480,507, src/cmd_tag.rs:393-396).
//
// HBM-bound by design: every text byte is read exactly once with coalesced 16-byte loads (read-only
// path, no L1 allocation, evict-first in L2), packed to 2-bit classes in registers, and one seed per
// D bases is probed in a blocked Bloom filter staged in shared memory (one LDS per seed). The rare
// survivors go through the L2-resident cuckoo seed table and an exact byte compare against the
// pattern bytes, which is what makes the result identical to the reference's byte matchers.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/merkurio_cuda.h"
#include "mk_codes.h"
#include "mk_tables.h"

// -DMK_DEBUG_CHECKS (merkurio_b200.build.build_cuda(debug_checks=True) -> libmerkurio_cuda_dbg.so): device-side
// asserts on every index the kernels compute for shared-memory queues, filters, tables and lists. compute-sanitizer
// is closed on the GPU pool this was developed on, so the parity suite is run against this build instead
// (profiles/sanitizer/README.md). Compiled out otherwise.
#ifdef MK_DEBUG_CHECKS
#include <cassert>
#define MK_ASSERT(c) assert(c)
#else
#define MK_ASSERT(c) ((void)0)
#endif

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
    uint32_t n_vec;      // 16-byte vectors that cover them
    uint32_t n_records;
    const unsigned long long* off;  // n_records + 1 unit offsets
    const uint32_t* lens;           // optional record lengths
    // tables
    const uint32_t* filter;
    const uint32_t* filter2;        // optional second-level bitmap (nullptr: none)
    uint32_t filter_log2_bits;      // global flavour
    uint32_t filter_blocks;         // shared-memory flavour: 64-bit blocks
    uint32_t filter32;              // shared-memory flavour: 1 = 32-bit blocks, 3 bits per key (filter_probe32)
    uint32_t filter2_log2_bits;
    uint32_t bucket_mask;
    const SeedSlot* slots;
    const uint32_t* postings;
    const uint8_t* pat_bytes;
    const uint32_t* pat_off;
    const uint32_t* tie_rank;
    uint32_t q;
    uint32_t short_shift;           // 32 - 2q: candidate code (16-base window) -> q-base seed code (0: the code is the seed code)
    uint32_t win_mask0, win_mask1;  // mk_scan_win: the q bases of a window that form the seed (ASCII: code bits; BAM4: the two words)
    uint32_t has_long;              // 1: patterns of the long group are keyed by the whole 16-base window
    int case_insensitive;
    uint32_t short_mask;            // candidate code -> short seed code: (code >> short_shift) & short_mask
    uint32_t pos_flags2;            // 1: bit 1 of a candidate position says "only the short key passed" (mk_scan_dual8)
    uint32_t direct_shift, direct_words;  // mk_scan_short: 32 - 2 * q1 and the size of the direct bitmap in 32-bit words
    uint32_t gate_mask, gate_val;   // alphabet gate (mk_scan_dual8): byte b with (b ^ gate_val) & gate_mask != 0 occurs in no pattern (per byte lane)
    // candidates: seeds that passed both filters, handed from the scan to the verify kernel
    uint2* cand;                    // {seed position / pos_mul, seed code}
    unsigned long long cand_capacity;
    unsigned long long* cand_count;
    uint32_t pos_mul;               // 16 for the stride-16 scan, 1 otherwise
    // outputs
    uint32_t* flags;                // 1 bit / record
    RawHit* hits;
    unsigned long long hit_capacity;
    unsigned long long* hit_count;
    int mode;
    uint32_t len_bits, tie_bits, pat_bits, max_len;
    uint32_t n_patterns, n_postings;  // bounds for the debug checks
    unsigned long long* cta_clock;    // MK_CTA_CLOCKS=1 (diagnostics): [2 * blockIdx.x] = globaltimer at CTA start, [+1] at its end
};

__device__ __forceinline__ void cta_clock_mark(const ScanParams& P, int which) {
    if (P.cta_clock && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.cta_clock[2 * blockIdx.x + which] = t;
        if (which == 0) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.cta_clock[2 * gridDim.x + blockIdx.x] = smid;
        }
    }
}

constexpr int kScanThreads = 1024;
constexpr int kScanWarps = kScanThreads / 32;

// First-level filter flavours
constexpr int kFilterSmem = 0;    // blocked Bloom in shared memory: 4 bits inside one 64-bit block
constexpr int kFilterGlobal = 1;  // L2-resident: plain 1-hash bitmap (stride 16) or dual-key 64-bit blocks (stride < 16)

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

// 32-byte load (sm_100: ld.global.v8.b32), same cache behaviour without a policy operand
__device__ __forceinline__ void ld_stream32(const uint4* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

// One probe of the first-level filter. kFilterSmem: `f` is the shared-memory copy and `lb` the
// number of 64-bit blocks; kFilterGlobal: `lb` is log2 of the bitmap size.
template <int FMODE>
__device__ __forceinline__ uint32_t filter_probe(const uint32_t* __restrict__ f, uint32_t code, uint32_t lb) {
    if (FMODE == kFilterSmem) {
        uint32_t h = code * MK_BLOOM_MUL;
        MK_ASSERT(__umulhi(h, lb) < lb);
        uint2 w = reinterpret_cast<const uint2*>(f)[__umulhi(h, lb)];
        uint32_t g = h * MK_BLOOM_MUL2;
        // shifts by a register use its low 5 bits (SHF.R.W), so the bit positions need no masking
        return (w.x >> ((g >> 12) & 31)) & (w.x >> ((g >> 17) & 31)) & (w.y >> ((g >> 22) & 31)) & (w.y >> (g >> 27)) & 1u;
    } else {
        uint32_t h = mk_hash_f1(code, lb);
        return (__ldg(f + (h >> 5)) >> (h & 31)) & 1u;
    }
}

// Shared-memory filter with 32-bit blocks and 3 bits per key: one LDS.32 per probe, half the bank traffic of
// the 64-bit flavour. Chosen by the table builder while it stays selective (small seed sets); `lb` counts
// 64-bit units as for the other flavour.
__device__ __forceinline__ uint32_t filter_probe32(const uint32_t* __restrict__ f, uint32_t code, uint32_t lb) {
    uint32_t h = code * MK_BLOOM_MUL;
    uint32_t w = f[__umulhi(h, 2u * lb)];
    uint32_t g = h * MK_BLOOM_MUL2;
    return (w >> ((g >> 17) & 31)) & (w >> ((g >> 22) & 31)) & (w >> (g >> 27)) & 1u;
}

// Dual-key probe of the L2-resident blocked filter (stride < 16): `win` is the 16-base window at the
// grid position; its q-base prefix selects the block, and the block is tested for the prefix (short
// group, result bit 0) and, if a long group exists, for the whole window (bit 1). One 8-byte load.
__device__ __forceinline__ uint32_t bloom_test(uint2 w, uint32_t g) {
    return (w.x >> ((g >> 12) & 31)) & (w.x >> ((g >> 17) & 31)) & (w.y >> ((g >> 22) & 31)) & (w.y >> (g >> 27)) & 1u;
}
__device__ __forceinline__ uint32_t dual_probe(const uint32_t* __restrict__ f, uint32_t win, uint32_t short_shift,
                                               uint32_t nblocks, uint32_t has_long) {
    uint32_t sc = win >> short_shift;
    MK_ASSERT(mk_dual_block(sc, nblocks) < nblocks);
    uint2 w = __ldg(reinterpret_cast<const uint2*>(f) + mk_dual_block(sc, nblocks));
    uint32_t pass = bloom_test(w, mk_dual_g_short(sc));           // bit 0: short key present
    if (has_long) pass |= bloom_test(w, mk_dual_g_long(win)) << 1;  // bit 1: long key present
    return pass;
}

// The record that contains unit position s: the last r with off[r] <= s (records of length 0 are
// skipped by construction). Runs in one lane for one hit, so every dependent load costs a full L2 /
// DRAM round trip: interpolate first (exact in one step for reads of uniform length), then bisect
// what is left.
__device__ __forceinline__ uint32_t find_record(const unsigned long long* __restrict__ off, uint32_t n, uint64_t s) {
    uint32_t lo = 0, hi = n;  // invariant: off[lo] <= s < off[hi]
    unsigned long long olo = off[0], ohi = off[n];
#ifndef MK_NO_INTERP
    for (int it = 0; it < 4 && hi - lo > 1 && ohi > olo; ++it) {
        double frac = (double)(s - olo) / (double)(ohi - olo);
        uint32_t g = lo + (uint32_t)(frac * (double)(hi - lo));
        if (g >= hi) g = hi - 1;
        unsigned long long og = off[g], og1 = off[g + 1];
        if (og <= s) {
            if (s < og1) return g;
            lo = g + 1; olo = og1;
        } else {
            hi = g; ohi = og;
        }
    }
#endif
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

// Eight bytes from any byte address: two aligned 8-byte loads and a funnel shift. Reads up to 15 bytes
// past p (the text and pattern buffers are padded for that).
__device__ __forceinline__ uint64_t ld8_any(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a & ~uintptr_t(7));
    const uint32_t sh = (uint32_t)(a & 7u) * 8u;
    unsigned long long lo = __ldg(q), hi = __ldg(q + 1);
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}
// mk_fold on eight bytes at once: 'A'..'Z' -> 'a'..'z', every other byte (also >= 0x80) unchanged
__device__ __forceinline__ uint64_t fold8(uint64_t c) {
    const uint64_t t = c & 0x7F7F7F7F7F7F7F7Full;
    const uint64_t ge_a = t + 0x3F3F3F3F3F3F3F3Full;  // bit 7 set iff t >= 'A'
    const uint64_t gt_z = t + 0x2525252525252525ull;  // bit 7 set iff t >  'Z'
    const uint64_t upper = ge_a & ~gt_z & ~c & 0x8080808080808080ull;
    return c | (upper >> 2);
}

// Second-level filter, inlined at the drain sites: one L2 load and a handful of registers, so that
// the out-of-line verify (whose call spills the prefetched tile registers) runs for real candidates
// only.
__device__ __forceinline__ bool second_level_pass(const ScanParams& P, uint32_t code) {
    if (!P.filter2) return true;
    uint32_t h = mk_hash_f2(code >> P.short_shift, P.filter2_log2_bits);
    return (__ldg(P.filter2 + (h >> 5)) >> (h & 31)) & 1u;
}

// Cuckoo lookup of (group, code): two 32-byte buckets. Returns the first posting or kEmptySlot.
__device__ __forceinline__ uint32_t seed_lookup(const ScanParams& P, uint32_t code, uint32_t group) {
    const uint32_t hk = mk_group_key(code, group);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t b = h == 0 ? mk_hash_b1(hk, P.bucket_mask) : mk_hash_b2(hk, P.bucket_mask);
        MK_ASSERT(b <= P.bucket_mask);
        const uint4* bp = reinterpret_cast<const uint4*>(P.slots + (size_t)b * kBucketSlots);
        uint4 lo = __ldg(bp), hi = __ldg(bp + 1);
        if (lo.x == code && lo.y != kEmptySlot && (lo.y >> 31) == group) return lo.y & ~kGroupBit;
        if (lo.z == code && lo.w != kEmptySlot && (lo.w >> 31) == group) return lo.w & ~kGroupBit;
        if (hi.x == code && hi.y != kEmptySlot && (hi.y >> 31) == group) return hi.y & ~kGroupBit;
        if (hi.z == code && hi.w != kEmptySlot && (hi.w >> 31) == group) return hi.w & ~kGroupBit;
    }
    return kEmptySlot;
}

// Hits found by one warp of the verify kernel are staged in shared memory and appended to the global
// list 32 or more at a time with ONE atomicAdd (at BASELINE cfg5 a million hits would otherwise queue
// up on the single list counter).
constexpr int kHitStage = 64;
struct HitSink {
    RawHit* buf;    // kHitStage entries of the warp
    uint32_t* cnt;  // hits staged (may exceed kHitStage: the excess went straight to the global list)
};

__device__ __forceinline__ void append_hit_global(const ScanParams& P, const RawHit& h) {
    cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(P.hit_count, (unsigned long long)g.size());
    base = g.shfl(base, 0);
    unsigned long long slot = base + g.thread_rank();
    if (slot < P.hit_capacity) P.hits[slot] = h;
}

// Called by the whole warp: move the staged hits to the global list.
__device__ __forceinline__ void sink_flush(const ScanParams& P, HitSink& sink, uint32_t lane) {
    __syncwarp();
    uint32_t n = *sink.cnt;
    if (n > kHitStage) n = kHitStage;
    if (n) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.hit_count, (unsigned long long)n);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (uint32_t k = lane; k < n; k += 32)
            if (base + k < P.hit_capacity) P.hits[base + k] = sink.buf[k];
    }
    __syncwarp();
    if (lane == 0) *sink.cnt = 0;
    __syncwarp();
}

// One posting (pattern pid owns the seed at its offset j) of a seed key found at base position `pos`:
// compare the pattern byte by byte. Runs in the verify kernel, one candidate per thread.
template <int ENC>
__device__ __forceinline__ void verify_one(const ScanParams& P, uint64_t pos, uint32_t pid, uint32_t j, HitSink& sink) {
    if (pos < j) return;
    MK_ASSERT(pid < P.n_patterns);  // (pos may lie past the text: the zero padding of a ragged last tile looks like poly-A and is dropped below)
    const uint8_t* text = reinterpret_cast<const uint8_t*>(P.text);
    const uint64_t s = pos - j;
    const uint32_t po = __ldg(P.pat_off + pid);
    const uint32_t L = __ldg(P.pat_off + pid + 1) - po;
    if (s + L > P.n_units) return;
    const uint8_t* pat = P.pat_bytes + po;
    bool eq = true;
    if (ENC == MK_ENC_ASCII) {
        // exact compare, eight bytes per step
        for (uint32_t k0 = 0; k0 < L && eq; k0 += 8) {
            uint64_t tx = ld8_any(text + s + k0), px = ld8_any(pat + k0);
            if (P.case_insensitive) tx = fold8(tx);
            uint64_t diff = tx ^ px;
            if (L - k0 < 8) diff &= (1ull << (8 * (L - k0))) - 1ull;
            eq = diff == 0;
        }
    } else {
        // nibble text against one-nibble-per-byte pattern codes, eight symbols per round trip
        for (uint32_t k0 = 0; k0 < L && eq; k0 += 8) {
            uint8_t tx[8], px[8];
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) {
                bool in = k0 + i < L;
                tx[i] = in ? text_symbol<ENC>(text, s + k0 + i, P.case_insensitive) : 0;
                px[i] = in ? __ldg(pat + k0 + i) : 0;
            }
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) eq = eq && (tx[i] == px[i]);
        }
    }
    if (!eq || s < P.off[0]) return;
    const uint32_t r = find_record(P.off, P.n_records, s);
    MK_ASSERT(r < P.n_records && P.off[r] <= s);
    const uint64_t rend = P.lens ? P.off[r] + P.lens[r] : P.off[r + 1];
    if (s + L > rend) return;
    // test before set: with few long records (chromosomes) every hit lands on the same word
    const uint32_t bit = 1u << (r & 31);
    if (!(__ldcg(P.flags + (r >> 5)) & bit)) atomicOr(P.flags + (r >> 5), bit);
    if (P.mode == MK_MODE_FLAG) return;
    RawHit hrec;
    if (P.mode == MK_MODE_ALL_HITS)
        hrec.key = ((((unsigned long long)(s + L) << P.len_bits) | (P.max_len - L)) << P.tie_bits) | __ldg(P.tie_rank + pid);
    else
        hrec.key = ((unsigned long long)r << P.pat_bits) | pid;
    hrec.record = r;
    hrec.pattern = pid;
    const uint32_t k = atomicAdd(sink.cnt, 1u);
    if (k < kHitStage) sink.buf[k] = hrec;
    else append_hit_global(P, hrec);
}

// All postings of a key: `first` is what seed_lookup returned (group bit stripped).
template <int ENC>
__device__ __forceinline__ void verify_postings(const ScanParams& P, uint64_t pos, uint32_t first, HitSink& sink) {
    if (first & kInlineBit) {
        verify_one<ENC>(P, pos, (first & ~kInlineBit) >> 4, first & 15u, sink);
        return;
    }
    for (uint32_t i = first;; ++i) {
        MK_ASSERT(i < P.n_postings);
        const uint32_t e = __ldg(P.postings + i);
        verify_one<ENC>(P, pos, e >> 5, (e >> 1) & 15u, sink);
        if (e & 1u) break;
    }
}

// A candidate: `code` is the seed code of the stride-16 scan, or the 16-base window of the other
// scans (its q-base prefix is the short key, the whole window the long key).
template <int ENC>
__device__ __forceinline__ void verify_seed(const ScanParams& P, uint64_t pos, uint32_t code, HitSink& sink) {
    bool long_only = false, short_only = false;
    if (P.pos_flags2) {  // mk_scan_dual8 (stride 8): bits 0..1 of the position say which single key passed, if only one did
        long_only = pos & 1u;
        short_only = pos & 2u;
        pos &= ~3ull;
    } else if (P.has_long) {  // bit 0 of the position is a flag of the scan (the stride is even)
        long_only = pos & 1u;
        pos &= ~1ull;
    }
    uint32_t f_long = (P.has_long && !short_only) ? seed_lookup(P, code, 1u) : kEmptySlot;
    uint32_t f_short = long_only ? kEmptySlot : seed_lookup(P, (code >> P.short_shift) & P.short_mask, 0u);
    if (f_long != kEmptySlot) verify_postings<ENC>(P, pos, f_long, sink);
    if (f_short != kEmptySlot) verify_postings<ENC>(P, pos, f_short, sink);
}

template <int FMODE>
__device__ __forceinline__ void stage_filter(const ScanParams& P, uint32_t* s_filter) {
    if (FMODE != kFilterSmem) return;
    const uint32_t n16 = P.filter_blocks / 2;  // uint4 count (the block count is even)
    const uint4* src = reinterpret_cast<const uint4*>(P.filter);
    uint4* dst = reinterpret_cast<uint4*>(s_filter);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// D == 16: the seed is the 16-base unit itself; no lane needs its neighbour.
// ASCII: one seed per 16-byte vector. BAM4: two seeds per vector.
//
// Work split: a tile is U rows of 512 contiguous bytes (lane l owns vector l of every row); warp w of
// the grid takes tiles w, w + W, w + 2W, ... The loop is unrolled twice over two register buffers so
// that the loads of the next tile are in flight while the current one is processed (2*U loads per
// lane) and no register copies are needed. Only full tiles run through the loop; the ragged last
// tile is handled once, with bounds checks, after it.
// ---------------------------------------------------------------------------------------------
// Candidate queue of one warp (shared memory). Seeds that pass the first-level filter are rare and
// scattered over the lanes; verifying them where they occur would stall the whole warp on two
// dependent L2 reads for one or two active lanes. Instead they are compacted (ballot + popc) into
// this queue and verified 32 at a time, one candidate per lane, with all lookups in flight together.
constexpr int kQueueCap = 64;
struct WarpQueue {
    uint2* slot;     // kQueueCap entries of {seed position / PM, seed code}; PM = 16 (stride-16 scan) or 1
    uint32_t count;  // warp-uniform
    uint32_t* cnt;   // shared-memory copy of `count`, the target of the lanes' slot-claiming atomics
};

// Lanes whose candidate also passes the second-level (L2-resident) filter append it to the global
// candidate list: one warp-aggregated atomic per drain.
__device__ __forceinline__ void emit_candidates(const ScanParams& P, bool mine, uint2 e, uint32_t lane) {
    uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
    if (m == 0) return;
    unsigned long long base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(P.cand_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
    if (mine && slot < P.cand_capacity) P.cand[slot] = e;
}

__device__ __forceinline__ void queue_drain32(const ScanParams& P, WarpQueue& wq, uint32_t lane) {
    MK_ASSERT(wq.count >= 32 && wq.count <= kQueueCap);
    wq.count -= 32;
    uint2 e = wq.slot[wq.count + lane];
    __syncwarp();
    emit_candidates(P, second_level_pass(P, e.y), e, lane);
}
__device__ __forceinline__ void queue_flush(const ScanParams& P, WarpQueue& wq, uint32_t lane) {
    __syncwarp();
    uint2 e = make_uint2(0, 0);
    bool mine = false;
    if (lane < wq.count) {
        e = wq.slot[lane];
        mine = second_level_pass(P, e.y);
    }
    emit_candidates(P, mine, e, lane);
    wq.count = 0;
    __syncwarp();
}
// push the lanes whose `mine` is set; unit (= position / PM) and code are per lane
__device__ __forceinline__ void queue_push(const ScanParams& P, WarpQueue& wq, uint32_t lane, bool mine, uint32_t unit,
                                           uint32_t code) {
    uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
    if (m) {
        MK_ASSERT(wq.count + __popc(m) <= kQueueCap);
        if (mine) wq.slot[wq.count + __popc(m & ((1u << lane) - 1u))] = make_uint2(unit, code);
        wq.count += __popc(m);
        __syncwarp();
        if (wq.count >= 32) queue_drain32(P, wq, lane);
    }
}

// v0 = index of the lane's first vector of the tile. V8: the lane owns pairs of adjacent vectors
// (32-byte loads, rows of 64 vectors); else single vectors (rows of 32 vectors).
template <int ENC, int FMODE, int U, bool V8, bool B32>
__device__ __forceinline__ void process_tile(const ScanParams& P, const uint32_t* __restrict__ filt, uint32_t lb,
                                             const uint4 (&v)[U], uint32_t v0, WarpQueue& wq, uint32_t lane) {
    constexpr int SPV = (ENC == MK_ENC_ASCII) ? 1 : 2;  // seeds (= units) per vector
    uint32_t code[U * SPV];
    uint32_t pass = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (ENC == MK_ENC_ASCII) {
            code[u] = mk_pack_ascii_perm(v[u].x, v[u].y, v[u].z, v[u].w);
        } else {
            code[2 * u] = mk_pack_bam_perm(v[u].x, v[u].y);
            code[2 * u + 1] = mk_pack_bam_perm(v[u].z, v[u].w);
        }
    }
#pragma unroll
    for (int k = 0; k < U * SPV; ++k) pass |= (B32 ? filter_probe32(filt, code[k], lb) : filter_probe<FMODE>(filt, code[k], lb)) << k;
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, __popc(pass));
    if (total == 0) return;
    if (wq.count + total <= kQueueCap) {
        // common case: the few lanes that hold candidates claim their slots with one shared-memory
        // atomic each; no per-seed warp votes
        if (pass) {
            uint32_t idx = atomicAdd(wq.cnt, __popc(pass));
            MK_ASSERT(idx + __popc(pass) <= kQueueCap);
#pragma unroll
            for (int k = 0; k < U * SPV; ++k) {
                const int u = k / SPV;
                const uint32_t vec = V8 ? v0 + (u / 2) * 64 + (u % 2) : v0 + u * 32;
                if ((pass >> k) & 1u) wq.slot[idx++] = make_uint2(vec * SPV + (k % SPV), code[k]);
            }
        }
        __syncwarp();
        wq.count += total;
        if (wq.count >= 32) {
            do queue_drain32(P, wq, lane); while (wq.count >= 32);
            if (lane == 0) *wq.cnt = wq.count;
            __syncwarp();
        }
    } else {
        // a burst of candidates that could overflow the queue: push seed by seed, draining on the way
#pragma unroll
        for (int k = 0; k < U * SPV; ++k) {
            const int u = k / SPV;
            const uint32_t vec = V8 ? v0 + (u / 2) * 64 + (u % 2) : v0 + u * 32;
            queue_push(P, wq, lane, (pass >> k) & 1u, vec * SPV + (k % SPV), code[k]);
        }
        if (lane == 0) *wq.cnt = wq.count;
        __syncwarp();
    }
}

template <int U, bool V8>
__device__ __forceinline__ void load_rows(const uint4* __restrict__ p, uint64_t pol, uint4 (&v)[U]) {
    if (V8) {
#pragma unroll
        for (int u = 0; u < U; u += 2) ld_stream32(p + (u / 2) * 64, v[u], v[u + 1]);
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_stream(p + u * 32, pol);
    }
}

template <int ENC, int FMODE, int U, int T, bool V8, bool B32 = false>
__global__ void __launch_bounds__(T, 1) mk_scan_d16(const __grid_constant__ ScanParams P) {
    static_assert(!V8 || U % 2 == 0, "32-byte loads need an even number of vectors per lane");
    constexpr int kScanWarps = T / 32;
    extern __shared__ __align__(16) uint32_t s_filter[];
    __shared__ uint2 s_queue[kScanWarps][kQueueCap];
    __shared__ uint32_t s_qcount[kScanWarps];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lane_vec = V8 ? lane * 2 : lane;
    const uint32_t lb = (FMODE == kFilterSmem) ? P.filter_blocks : P.filter_log2_bits;
    const uint64_t pol = V8 ? 0 : make_evict_first_policy();
    const uint32_t nwarps = gridDim.x * kScanWarps;
    const uint32_t full_tiles = P.n_vec / (U * 32);
    cta_clock_mark(P, 0);
    WarpQueue wq{s_queue[threadIdx.x >> 5], 0, &s_qcount[threadIdx.x >> 5]};
    if (lane == 0) *wq.cnt = 0;
    __syncwarp();
    uint32_t t = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
    const uint32_t warp0 = t;
    const size_t stride = (size_t)nwarps * (U * 32);
    const uint4* p = P.text + (size_t)t * (U * 32) + lane_vec;

    uint4 a[U], b[U];
    if (t < full_tiles) load_rows<U, V8>(p, pol, a);  // in flight while the filter is staged
    stage_filter<FMODE>(P, s_filter);
    const uint32_t* __restrict__ filt = (FMODE == kFilterSmem) ? s_filter : P.filter;

    while (t < full_tiles) {
        uint32_t tn = t + nwarps;
        if (tn < full_tiles) load_rows<U, V8>(p + stride, pol, b);
        process_tile<ENC, FMODE, U, V8, B32>(P, filt, lb, a, t * (U * 32) + lane_vec, wq, lane);
        t = tn;
        p += stride;
        if (t >= full_tiles) break;
        tn = t + nwarps;
        if (tn < full_tiles) load_rows<U, V8>(p + stride, pol, a);
        process_tile<ENC, FMODE, U, V8, B32>(P, filt, lb, b, t * (U * 32) + lane_vec, wq, lane);
        t = tn;
        p += stride;
    }
    // ragged last tile (16-byte loads with bounds checks)
    if (P.n_vec % (U * 32) != 0 && warp0 == full_tiles % nwarps) {
        const uint64_t pol16 = make_evict_first_policy();
        const uint32_t v0 = full_tiles * (U * 32) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (v0 + u * 32 < P.n_vec) ? ld_stream(P.text + v0 + u * 32, pol16) : make_uint4(0, 0, 0, 0);
        process_tile<ENC, FMODE, U, false, B32>(P, filt, lb, a, v0, wq, lane);
    }
    queue_flush(P, wq, lane);
    if (P.cta_clock) { __syncthreads(); cta_clock_mark(P, 1); }
}

// ---------------------------------------------------------------------------------------------
// Experiment (MK_TMA=1, not the default — measured slower, profiles/r2_tma_experiment.txt): the stride-16 scan with its
// tiles staged through shared memory by the bulk-copy engine (TMA: cp.async.bulk.shared::cluster.global, completion on
// an mbarrier) instead of 16-byte register loads. One lane per warp issues one bulk copy per tile (U x 512 contiguous
// bytes) into one of the warp's two stages; the lanes wait on the stage's mbarrier and fetch their vectors with
// LDS.128. The register double buffer goes away (fewer registers, no load instructions in the lanes), but every text
// byte now crosses shared memory twice (written by the copy engine, read by LDS) in a kernel whose shared-memory
// filter probes already keep L1TEX busy, and the stages take shared memory from the filter's neighbour, the L1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MK_DONE;\n"
        "bra MK_WAIT;\n"
        "MK_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

template <int ENC, int U, int T, bool B32>
__global__ void __launch_bounds__(T, 1) mk_scan_d16_tma(const __grid_constant__ ScanParams P) {
    constexpr int kWarps = T / 32;
    constexpr uint32_t kTileBytes = U * 512;
    extern __shared__ __align__(128) uint32_t s_dyn[];  // [filter | kWarps x 2 stages x kTileBytes]
    __shared__ uint2 s_queue[kWarps][kQueueCap];
    __shared__ uint32_t s_qcount[kWarps];
    __shared__ __align__(8) uint64_t s_bar[kWarps][2];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lb = P.filter_blocks;
    const uint32_t filter_bytes = (P.filter_blocks * 8u + 127u) & ~127u;
    uint8_t* stage0 = reinterpret_cast<uint8_t*>(s_dyn) + filter_bytes + (size_t)w * 2 * kTileBytes;
    const uint64_t pol = make_evict_first_policy();
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t full_tiles = P.n_vec / (U * 32);
    WarpQueue wq{s_queue[w], 0, &s_qcount[w]};
    if (lane == 0) {
        *wq.cnt = 0;
        mbar_init(&s_bar[w][0], 1);
        mbar_init(&s_bar[w][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t t = blockIdx.x * kWarps + w;
    const uint32_t warp0 = t;
    const uint8_t* text = reinterpret_cast<const uint8_t*>(P.text);
    auto issue = [&](uint32_t tile, uint32_t st) {  // lane 0: one bulk copy of the tile into stage st
        mbar_expect_tx(&s_bar[w][st], kTileBytes);
        bulk_g2s(stage0 + st * kTileBytes, text + (size_t)tile * kTileBytes, kTileBytes, &s_bar[w][st], pol);
    };
    if (lane == 0) {
        if (t < full_tiles) issue(t, 0);
        if (t + nwarps < full_tiles) issue(t + nwarps, 1);
    }
    stage_filter<kFilterSmem>(P, s_dyn);
    const uint32_t* __restrict__ filt = s_dyn;

    uint32_t st = 0, phase = 0;
    uint4 a[U];
    while (t < full_tiles) {
        mbar_wait(&s_bar[w][st], phase);
        const uint4* sv = reinterpret_cast<const uint4*>(stage0 + st * kTileBytes) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = sv[u * 32];
        __syncwarp();  // every lane has its vectors: the stage may be refilled
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t tn2 = t + 2 * nwarps;
            if (tn2 < full_tiles) issue(tn2, st);
        }
        process_tile<ENC, kFilterSmem, U, false, B32>(P, filt, lb, a, t * (U * 32) + lane, wq, lane);
        t += nwarps;
        st ^= 1;
        if (st == 0) phase ^= 1;
    }
    // ragged last tile (16-byte loads with bounds checks)
    if (P.n_vec % (U * 32) != 0 && warp0 == full_tiles % nwarps) {
        const uint32_t v0 = full_tiles * (U * 32) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (v0 + u * 32 < P.n_vec) ? ld_stream(P.text + v0 + u * 32, pol) : make_uint4(0, 0, 0, 0);
        process_tile<ENC, kFilterSmem, U, false, B32>(P, filt, lb, a, v0, wq, lane);
    }
    queue_flush(P, wq, lane);
}

// ---------------------------------------------------------------------------------------------
// D == 8 or 4 with the shared-memory filter: same tile loop as the stride-16 scan. A seed is the
// 16-base window that starts D, 2D, ... bases into a unit; in the permuted packing (base i of a unit
// in field i / 4 of byte lane i % 4) that window is two shifts and a select of (unit code, next unit
// code) — no ordered packing, no funnel shifts — and for BAM4 it is a pair of adjacent words. Seeds
// shorter than 16 bases (k < D + 15) mask the window's tail. The unit after a lane's vector comes
// from the next lane, from lane 0 of the next row, or (last row, lane 31) from one extra 16-byte
// load of the first vector of the following tile.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void push_tile_candidates(const ScanParams& P, WarpQueue& wq, uint32_t lane, uint32_t pass,
                                                     const uint32_t (&code)[K], uint32_t unit0, uint32_t unit_row_stride, int per_row) {
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, __popc(pass));
    if (total == 0) return;
    if (wq.count + total <= kQueueCap) {
        if (pass) {
            uint32_t idx = atomicAdd(wq.cnt, __popc(pass));
            MK_ASSERT(idx + __popc(pass) <= kQueueCap);
#pragma unroll
            for (int k = 0; k < K; ++k)
                if ((pass >> k) & 1u) wq.slot[idx++] = make_uint2(unit0 + (k / per_row) * unit_row_stride + (k % per_row), code[k]);
        }
        __syncwarp();
        wq.count += total;
        if (wq.count >= 32) {
            do queue_drain32(P, wq, lane); while (wq.count >= 32);
            if (lane == 0) *wq.cnt = wq.count;
            __syncwarp();
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
            queue_push(P, wq, lane, (pass >> k) & 1u, unit0 + (k / per_row) * unit_row_stride + (k % per_row), code[k]);
        if (lane == 0) *wq.cnt = wq.count;
        __syncwarp();
    }
}

template <int ENC, int D, int U, bool B32>
__device__ __forceinline__ void process_tile_win(const ScanParams& P, const uint32_t* __restrict__ filt, uint32_t lb, const uint4 (&v)[U],
                                                 const uint4& halo, uint32_t v0, WarpQueue& wq, uint32_t lane) {
    constexpr int PPV = (ENC == MK_ENC_ASCII ? 16 : 32) / D;  // probes per vector
    constexpr int K = U * PPV;
    static_assert(K <= 32, "the pass mask of a tile is one 32-bit word");
    uint32_t win[K];
    uint32_t pass = 0;
    if (ENC == MK_ENC_ASCII) {
        const uint32_t mask = P.win_mask0;
        uint32_t c[U + 1];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = mk_pack_ascii_perm(v[u].x, v[u].y, v[u].z, v[u].w);
        c[U] = mk_pack_ascii_perm(halo.x, halo.y, halo.z, halo.w);  // meaningful in lane 31 only
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t from_next_lane = __shfl_down_sync(0xFFFFFFFFu, c[u], 1);
            const uint32_t from_next_row = (u + 1 < U) ? __shfl_sync(0xFFFFFFFFu, c[u + 1], 0) : c[U];
            const uint32_t succ = (lane == 31) ? from_next_row : from_next_lane;
#pragma unroll
            for (int m = 0; m < PPV; ++m) {
                const int sh = 2 * (m * D / 4);  // the window starts m * D bases into the unit
                const uint32_t low = 0x01010101u * (0xFFu >> sh);
                const uint32_t w = (sh == 0) ? c[u] : (((c[u] >> sh) & low) | ((succ << (8 - sh)) & ~low));
                win[u * PPV + m] = w & mask;
            }
        }
    } else {
        const uint32_t m0 = P.win_mask0, m1 = P.win_mask1;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t from_next_lane = __shfl_down_sync(0xFFFFFFFFu, v[u].x, 1);
            const uint32_t from_next_row = (u + 1 < U) ? __shfl_sync(0xFFFFFFFFu, v[u + 1].x, 0) : halo.x;
            const uint32_t nx = (lane == 31) ? from_next_row : from_next_lane;
            static_assert(ENC == MK_ENC_ASCII || D == 8, "BAM4 windows are word aligned for a stride of 8 only");
            win[u * PPV + 0] = mk_pack_bam_perm(v[u].x & m0, v[u].y & m1);
            win[u * PPV + 1] = mk_pack_bam_perm(v[u].y & m0, v[u].z & m1);
            win[u * PPV + 2] = mk_pack_bam_perm(v[u].z & m0, v[u].w & m1);
            win[u * PPV + 3] = mk_pack_bam_perm(v[u].w & m0, nx & m1);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) pass |= (B32 ? filter_probe32(filt, win[k], lb) : filter_probe<kFilterSmem>(filt, win[k], lb)) << k;
    // candidate position / D: vector index * PPV + window index
    push_tile_candidates<K>(P, wq, lane, pass, win, v0 * PPV, 32 * PPV, PPV);
}

template <int ENC, int D, int U, int T, bool B32 = false>
__global__ void __launch_bounds__(T, 1) mk_scan_win(const __grid_constant__ ScanParams P) {
    constexpr int kWarps = T / 32;
    extern __shared__ __align__(16) uint32_t s_filter[];
    __shared__ uint2 s_queue[kWarps][kQueueCap];
    __shared__ uint32_t s_qcount[kWarps];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lb = P.filter_blocks;
    const uint64_t pol = make_evict_first_policy();
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t full_tiles = P.n_vec / (U * 32);
    WarpQueue wq{s_queue[threadIdx.x >> 5], 0, &s_qcount[threadIdx.x >> 5]};
    if (lane == 0) *wq.cnt = 0;
    __syncwarp();
    uint32_t t = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const uint32_t warp0 = t;
    const size_t stride = (size_t)nwarps * (U * 32);
    const uint4* p = P.text + (size_t)t * (U * 32) + lane;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    // first vector of the tile after tile `tt` (lane 31 only; zero past the end of the text)
    auto load_halo = [&](uint32_t tt) {
        const uint64_t hv = ((uint64_t)tt + 1) * (U * 32);
        return (lane == 31 && hv < P.n_vec) ? ld_stream(P.text + hv, pol) : zero;
    };

    uint4 a[U], b[U], ha = zero, hb = zero;
    if (t < full_tiles) { load_rows<U, false>(p, pol, a); ha = load_halo(t); }
    stage_filter<kFilterSmem>(P, s_filter);
    const uint32_t* __restrict__ filt = s_filter;

    while (t < full_tiles) {
        uint32_t tn = t + nwarps;
        if (tn < full_tiles) { load_rows<U, false>(p + stride, pol, b); hb = load_halo(tn); }
        process_tile_win<ENC, D, U, B32>(P, filt, lb, a, ha, t * (U * 32) + lane, wq, lane);
        t = tn;
        p += stride;
        if (t >= full_tiles) break;
        tn = t + nwarps;
        if (tn < full_tiles) { load_rows<U, false>(p + stride, pol, a); ha = load_halo(tn); }
        process_tile_win<ENC, D, U, B32>(P, filt, lb, b, hb, t * (U * 32) + lane, wq, lane);
        t = tn;
        p += stride;
    }
    // ragged last tile (bounds-checked loads; nothing follows it)
    if (P.n_vec % (U * 32) != 0 && warp0 == full_tiles % nwarps) {
        const uint32_t v0 = full_tiles * (U * 32) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (v0 + u * 32 < P.n_vec) ? ld_stream(P.text + v0 + u * 32, pol) : zero;
        process_tile_win<ENC, D, U, B32>(P, filt, lb, a, zero, v0, wq, lane);
    }
    queue_flush(P, wq, lane);
}

// ---------------------------------------------------------------------------------------------
// D < 16: ordered unit codes; a seed at offset o of a unit is a funnel shift of (unit, next unit).
// The next unit lives in the next lane (shuffle), in lane 0 of the next row, or in the first
// vector of the next tile (one extra load by lane 0).
// ---------------------------------------------------------------------------------------------
template <int ENC, int D, int FMODE, int U>
__global__ void __launch_bounds__(kScanThreads, 1) mk_scan_ord(const __grid_constant__ ScanParams P) {
    extern __shared__ __align__(16) uint32_t s_filter[];
    __shared__ uint2 s_queue[kScanWarps][kQueueCap];
    stage_filter<FMODE>(P, s_filter);
    const uint32_t* __restrict__ filt = (FMODE == kFilterSmem) ? s_filter : P.filter;
    const uint32_t lane = threadIdx.x & 31;
    WarpQueue wq{s_queue[threadIdx.x >> 5], 0, nullptr};
    const uint32_t lb = P.filter_blocks, sshift = P.short_shift, has_long = P.has_long;
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
                    // the 16-base window at the grid position: its q-base prefix is the short key
                    uint32_t win = mk_win_ord(cur, nxt, o);
                    uint32_t hit = (FMODE == kFilterSmem) ? filter_probe<kFilterSmem>(filt, win >> sshift, lb)
                                                          : dual_probe(filt, win, sshift, lb, has_long);
                    bool pass = hit && base + o < P.n_units;
                    // Positions fit 32 bits: the engine refuses batches of 2^32 bases or more on this path.
                    // They are multiples of the stride (>= 2 whenever a long group exists), so bit 0 can say
                    // "only the long key passed", which saves the verify kernel the short-key lookup.
                    queue_push(P, wq, lane, pass, (uint32_t)(base + o) | ((FMODE != kFilterSmem && hit == 2u) ? 1u : 0u), win);
                }
            }
        }
    }
    queue_flush(P, wq, lane);
}

// ---------------------------------------------------------------------------------------------
// D == 8, ASCII, L2-resident dual-key filter: query sets far too large for shared memory (BASELINE cfg5: a
// million queries of 21..63 bases, 8 million seeds). Two things bound this kernel, and it is written around both:
//   * gathers: every probe is a lane-divergent 8-byte load, and an SM retires about one of those per cycle whatever
//     the path (global, texture or distributed shared memory: scripts/micro/gather_bench.cu,
//     profiles/r2_gather_bench.txt) against ten per cycle from its own shared memory, which 8 million keys do not
//     fit. The alphabet gate removes the probes that cannot succeed: a window whose first q bytes hold a byte that
//     occurs in no pattern (lower-case soft-masked spans against an upper-case k-mer list, N runs, ...) cannot
//     start a seed — exact, tested on the raw bytes, one OR-chain per tile in the common all-clean case.
//   * integer instructions (the first version ran at 84 % of the ALU pipe: profiles/r2_cfg5_experiments/): keys in
//     the permuted window layout (8 instead of 22 operations per packed unit, a window is two shifts and a select),
//     3-bit filter entries addressed by the top bits of one product (a test is six shifts and an AND), multiplies
//     instead of shift-or wherever the two are the same (they issue on the other pipe), tiles without a live window
//     left after the gate are dropped before anything is packed.
// Same tile loop as the other scans (two tiles in flight per lane), all gathers of a tile issued before the first
// one is consumed, candidates queued with one shared-memory atomic per lane that holds any.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dual3_test(uint2 w, uint32_t g) {  // bit 0: all three bits of the key are set
    return __funnelshift_r(w.x, 0u, g >> 27) & __funnelshift_r(w.x, 0u, g >> 22) & __funnelshift_r(w.y, 0u, g >> 17);
}

template <int U, bool GATE>
__device__ __forceinline__ void process_tile_dual8(const ScanParams& P, const uint4 (&v)[U], const uint4& halo, uint32_t v0,
                                                   WarpQueue& wq, uint32_t lane) {
    constexpr int K = 2 * U;
    const uint32_t nblocks = P.filter_blocks, has_long = P.has_long, smask = P.win_mask0;
    // gate: bit k of `dead` = window k holds a byte no pattern contains within its first 12 (<= q) bytes
    uint32_t dead = 0;
    if (GATE) {
        const uint32_t M = P.gate_mask, V = P.gate_val;
        uint32_t acc = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) acc |= (v[u].x ^ V) | (v[u].y ^ V) | (v[u].z ^ V) | (v[u].w ^ V);
        if (lane == 31) acc |= (halo.x ^ V);  // the bytes the last window of the tile reaches into
        if (__any_sync(0xFFFFFFFFu, (acc & M) != 0)) {
            // rare: a span of foreign bytes (soft-masked, N, ...) starts, ends or runs through this tile
            uint32_t head_votes[U + 1], tail_bad = 0;
#pragma unroll
            for (int u = 0; u <= U; ++u) {
                const uint4& x = (u < U) ? v[u] : halo;
                const bool head = ((x.x ^ V) & M) != 0;  // bytes 0..3: what the window that starts 8 bytes earlier sees of this vector
                if (u < U) {
                    head_votes[u] = __ballot_sync(0xFFFFFFFFu, head);
                    if ((((x.x ^ V) | (x.y ^ V) | (x.z ^ V)) & M) != 0) dead |= 1u << (2 * u);  // bytes 0..11
                    if ((((x.z ^ V) | (x.w ^ V)) & M) != 0) tail_bad |= 1u << u;                // bytes 8..15
                } else {
                    head_votes[u] = head ? 1u : 0u;  // lane 31's halo vector
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t next_head = (lane == 31) ? (head_votes[u + 1] & 1u) : ((head_votes[u] >> (lane + 1)) & 1u);
                if (((tail_bad >> u) & 1u) | next_head) dead |= 2u << (2 * u);
            }
            if (__all_sync(0xFFFFFFFFu, dead == (1u << K) - 1u)) return;  // nothing in this tile can match
        }
    }
    uint32_t c[U + 1];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = mk_pack_ascii_perm(v[u].x, v[u].y, v[u].z, v[u].w);
    c[U] = mk_pack_ascii_perm(halo.x, halo.y, halo.z, halo.w);  // meaningful in lane 31 only
    uint32_t win[K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t from_next_lane = __shfl_down_sync(0xFFFFFFFFu, c[u], 1);
        const uint32_t from_next_row = (u + 1 < U) ? __shfl_sync(0xFFFFFFFFu, c[u + 1], 0) : c[U];
        const uint32_t succ = (lane == 31) ? from_next_row : from_next_lane;
        win[2 * u] = c[u];
        win[2 * u + 1] = ((c[u] >> 4) & 0x0F0F0F0Fu) | ((succ << 4) & 0xF0F0F0F0u);  // the window 8 bases into the unit
    }
    // all gathers of the tile, then the tests
    uint2 blk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        blk[k] = make_uint2(0, 0);
        MK_ASSERT(mk_dual_block(win[k] & smask, nblocks) < nblocks);
        if (!GATE || !((dead >> k) & 1u)) blk[k] = __ldg(reinterpret_cast<const uint2*>(P.filter) + mk_dual_block(win[k] & smask, nblocks));
    }
    uint32_t pass = 0, only = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t sh = dual3_test(blk[k], mk_dual3_g_short(win[k] & smask));
        const uint32_t lg = has_long ? dual3_test(blk[k], mk_dual3_g_long(win[k])) : 0u;
        pass += ((sh | lg) & 1u) * (1u << k);        // disjoint bits: a multiply-add is an OR (and issues on the other pipe)
        only += ((lg ^ sh) & 1u) * ((sh & 1u) + 1u) * (1u << (2 * k));  // 2-bit field k: 1 = only the long key passed, 2 = only the short key
    }
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, __popc(pass));
    if (total == 0) return;
    // candidate = {base position | which single key passed (bits 0..1), 16-base window}; positions are multiples of 8
    const uint32_t pos0 = v0 * 16u;
    if (wq.count + total <= kQueueCap) {
        if (pass) {
            uint32_t idx = atomicAdd(wq.cnt, __popc(pass));
            MK_ASSERT(idx + __popc(pass) <= kQueueCap);
#pragma unroll
            for (int k = 0; k < K; ++k)
                if ((pass >> k) & 1u) wq.slot[idx++] = make_uint2((pos0 + (k / 2) * 512u + (k % 2) * 8u) | ((only >> (2 * k)) & 3u), win[k]);
        }
        __syncwarp();
        wq.count += total;
        if (wq.count >= 32) {
            do queue_drain32(P, wq, lane); while (wq.count >= 32);
            if (lane == 0) *wq.cnt = wq.count;
            __syncwarp();
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
            queue_push(P, wq, lane, (pass >> k) & 1u, (pos0 + (k / 2) * 512u + (k % 2) * 8u) | ((only >> (2 * k)) & 3u), win[k]);
        if (lane == 0) *wq.cnt = wq.count;
        __syncwarp();
    }
}

template <int U, int T, bool GATE>
__global__ void __launch_bounds__(T, 1) mk_scan_dual8(const __grid_constant__ ScanParams P) {
    constexpr int kWarps = T / 32;
    __shared__ uint2 s_queue[kWarps][kQueueCap];
    __shared__ uint32_t s_qcount[kWarps];
    cta_clock_mark(P, 0);
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t pol = make_evict_first_policy();
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t full_tiles = P.n_vec / (U * 32);
    WarpQueue wq{s_queue[threadIdx.x >> 5], 0, &s_qcount[threadIdx.x >> 5]};
    if (lane == 0) *wq.cnt = 0;
    __syncwarp();
    // Adjacent tiles go to different CTAs (tile t -> CTA t mod grid): the gate makes soft-masked and N spans (kilobases =
    // tens to hundreds of tiles) nearly free, and spread over the CTAs they leave the SMs evenly loaded; with adjacent
    // tiles in one CTA the median CTA ended 50 us (5 %) before the last one on BASELINE cfg5.
    uint32_t t = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    const uint32_t warp0 = t;
    const size_t stride = (size_t)nwarps * (U * 32);
    const uint4* p = P.text + (size_t)t * (U * 32) + lane;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    auto load_halo = [&](uint32_t tt) {  // first vector of the tile after tile `tt` (lane 31 only; zero past the end of the text)
        const uint64_t hv = ((uint64_t)tt + 1) * (U * 32);
        return (lane == 31 && hv < P.n_vec) ? ld_stream(P.text + hv, pol) : zero;
    };

    uint4 a[U], b[U], ha = zero, hb = zero;
    if (t < full_tiles) { load_rows<U, false>(p, pol, a); ha = load_halo(t); }
    while (t < full_tiles) {
        uint32_t tn = t + nwarps;
        if (tn < full_tiles) { load_rows<U, false>(p + stride, pol, b); hb = load_halo(tn); }
        process_tile_dual8<U, GATE>(P, a, ha, t * (U * 32) + lane, wq, lane);
        t = tn;
        p += stride;
        if (t >= full_tiles) break;
        tn = t + nwarps;
        if (tn < full_tiles) { load_rows<U, false>(p + stride, pol, a); ha = load_halo(tn); }
        process_tile_dual8<U, GATE>(P, b, hb, t * (U * 32) + lane, wq, lane);
        t = tn;
        p += stride;
    }
    // ragged last tile (bounds-checked loads; nothing follows it)
    if (P.n_vec % (U * 32) != 0 && warp0 == full_tiles % nwarps) {
        const uint32_t v0 = full_tiles * (U * 32) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (v0 + u * 32 < P.n_vec) ? ld_stream(P.text + v0 + u * 32, pol) : zero;
        process_tile_dual8<U, GATE>(P, a, zero, v0, wq, lane);
    }
    queue_flush(P, wq, lane);
    if (P.cta_clock) { __syncthreads(); cta_clock_mark(P, 1); }
}

// ---------------------------------------------------------------------------------------------
// D == 1 or 2, ASCII, small query sets (shortest pattern below 15 bases: every or every second base starts a seed).
// With 8-16 probes per 16 bases the scan is bound by its instructions and by shared-memory probes, not by HBM, so the
// first level is the cheapest test there is: a DIRECT bitmap over the first q1 = min(q, 10) bases of a seed (4^10 bits
// = 128 KiB at most), indexed by the ordered 2-bit code itself — one shift, one LDS.32 and a bit test per position,
// no hashing (the blocked Bloom filter of mk_scan_ord costs about 16 instructions per probe). For q <= 10 that is
// exact set membership of the seed classes; for q = 11..13 the full seed is tested by the second-level (L2-resident)
// bitmap on the way out of the candidate queue, which only the ~1 % of positions that pass the prefix test reach
// (the builder keeps the Bloom filter when more than 1 % of the prefixes are occupied). Same double-buffered tile
// loop and batched queue insertion as the other scans (mk_scan_ord loads, probes and votes one probe at a time).
// ---------------------------------------------------------------------------------------------
template <int D, int U>
__device__ __forceinline__ void process_tile_short(const ScanParams& P, const uint32_t* __restrict__ bitmap, const uint4 (&v)[U],
                                                   const uint4& halo, uint32_t v0, WarpQueue& wq, uint32_t lane) {
    constexpr int PPV = 16 / D;  // probes per vector
    const uint32_t pshift = P.direct_shift;  // 32 - 2 * q1: window code -> index of its q1-base prefix
    uint32_t c[U + 1];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = mk_pack_ascii_ord(v[u].x, v[u].y, v[u].z, v[u].w);
    c[U] = mk_pack_ascii_ord(halo.x, halo.y, halo.z, halo.w);  // meaningful in lane 31 only
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t from_next_lane = __shfl_down_sync(0xFFFFFFFFu, c[u], 1);
        const uint32_t from_next_row = (u + 1 < U) ? __shfl_sync(0xFFFFFFFFu, c[u + 1], 0) : c[U];
        const uint32_t succ = (lane == 31) ? from_next_row : from_next_lane;
        uint32_t pass = 0;
#pragma unroll
        for (int k = 0; k < PPV; ++k) {
            const uint32_t idx = __funnelshift_l(succ, c[u], 2 * k * D) >> pshift;  // (k == 0: c[u] itself)
            pass += (__funnelshift_r(bitmap[idx >> 5], 0u, idx) & 1u) * (1u << k);
        }
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, __popc(pass));
        if (total == 0) continue;
        // candidate = {base position, 16-base window at that position}
        const uint32_t pos0 = (v0 + u * 32) * 16u;
        if (wq.count + total <= kQueueCap) {
            if (pass) {
                uint32_t idx = atomicAdd(wq.cnt, __popc(pass));
                MK_ASSERT(idx + __popc(pass) <= kQueueCap);
#pragma unroll
                for (int k = 0; k < PPV; ++k)
                    if ((pass >> k) & 1u) wq.slot[idx++] = make_uint2(pos0 + k * D, __funnelshift_l(succ, c[u], 2 * k * D));
            }
            __syncwarp();
            wq.count += total;
            if (wq.count >= 32) {
                do queue_drain32(P, wq, lane); while (wq.count >= 32);
                if (lane == 0) *wq.cnt = wq.count;
                __syncwarp();
            }
        } else {  // a burst (poly-A text against a poly-A pattern, ...): probe by probe, draining on the way
#pragma unroll
            for (int k = 0; k < PPV; ++k)
                queue_push(P, wq, lane, (pass >> k) & 1u, pos0 + k * D, __funnelshift_l(succ, c[u], 2 * k * D));
            if (lane == 0) *wq.cnt = wq.count;
            __syncwarp();
        }
    }
}

template <int D, int U, int T>
__global__ void __launch_bounds__(T, 1) mk_scan_short(const __grid_constant__ ScanParams P) {
    constexpr int kWarps = T / 32;
    extern __shared__ __align__(16) uint32_t s_bitmap[];
    __shared__ uint2 s_queue[kWarps][kQueueCap];
    __shared__ uint32_t s_qcount[kWarps];
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t pol = make_evict_first_policy();
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t full_tiles = P.n_vec / (U * 32);
    WarpQueue wq{s_queue[threadIdx.x >> 5], 0, &s_qcount[threadIdx.x >> 5]};
    if (lane == 0) *wq.cnt = 0;
    __syncwarp();
    uint32_t t = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const uint32_t warp0 = t;
    const size_t stride = (size_t)nwarps * (U * 32);
    const uint4* p = P.text + (size_t)t * (U * 32) + lane;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    auto load_halo = [&](uint32_t tt) {  // first vector of the tile after tile `tt` (lane 31 only; zero past the end of the text)
        const uint64_t hv = ((uint64_t)tt + 1) * (U * 32);
        return (lane == 31 && hv < P.n_vec) ? ld_stream(P.text + hv, pol) : zero;
    };
    uint4 a[U], b[U], ha = zero, hb = zero;
    if (t < full_tiles) { load_rows<U, false>(p, pol, a); ha = load_halo(t); }
    {   // in flight meanwhile: the bitmap into shared memory (direct_words is a multiple of 4)
        const uint4* src = reinterpret_cast<const uint4*>(P.filter);
        uint4* dst = reinterpret_cast<uint4*>(s_bitmap);
        for (uint32_t i = threadIdx.x; i < P.direct_words / 4; i += T) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    while (t < full_tiles) {
        uint32_t tn = t + nwarps;
        if (tn < full_tiles) { load_rows<U, false>(p + stride, pol, b); hb = load_halo(tn); }
        process_tile_short<D, U>(P, s_bitmap, a, ha, t * (U * 32) + lane, wq, lane);
        t = tn;
        p += stride;
        if (t >= full_tiles) break;
        tn = t + nwarps;
        if (tn < full_tiles) { load_rows<U, false>(p + stride, pol, a); ha = load_halo(tn); }
        process_tile_short<D, U>(P, s_bitmap, b, hb, t * (U * 32) + lane, wq, lane);
        t = tn;
        p += stride;
    }
    // ragged last tile (bounds-checked loads; nothing follows it)
    if (P.n_vec % (U * 32) != 0 && warp0 == full_tiles % nwarps) {
        const uint32_t v0 = full_tiles * (U * 32) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (v0 + u * 32 < P.n_vec) ? ld_stream(P.text + v0 + u * 32, pol) : zero;
        process_tile_short<D, U>(P, s_bitmap, a, zero, v0, wq, lane);
    }
    queue_flush(P, wq, lane);
}

// ---------------------------------------------------------------------------------------------
// Verification of the candidate list: one candidate per thread, so the dependent lookups (cuckoo
// bucket, postings, pattern bytes, record offsets) of thousands of candidates are in flight together
// instead of stalling a streaming warp. The list length is read from device memory.
// ---------------------------------------------------------------------------------------------
template <int ENC, int kVerifyThreads>
__global__ void __launch_bounds__(kVerifyThreads) mk_verify_candidates(const __grid_constant__ ScanParams P) {
    __shared__ RawHit s_hits[kVerifyThreads / 32][kHitStage];
    __shared__ uint32_t s_cnt[kVerifyThreads / 32];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    HitSink sink{s_hits[w], &s_cnt[w]};
    if (lane == 0) s_cnt[w] = 0;
    __syncwarp();
    unsigned long long n = *P.cand_count;
    if (n > P.cand_capacity) n = P.cand_capacity;
    const unsigned long long stride = (unsigned long long)gridDim.x * kVerifyThreads;
    // warp-uniform loop: all lanes of a warp stay in it together so that the staged hits can be flushed
    for (unsigned long long base = (unsigned long long)blockIdx.x * kVerifyThreads + w * 32; base < n; base += stride) {
        unsigned long long i = base + lane;
        if (i < n) {
            MK_ASSERT(i < P.cand_capacity);
            uint2 e = P.cand[i];
            verify_seed<ENC>(P, (uint64_t)e.x * P.pos_mul, e.y, sink);
        }
        __syncwarp();
        if (*sink.cnt >= 32) sink_flush(P, sink, lane);
    }
    sink_flush(P, sink, lane);
}

}  // namespace mk
