// 2-bit seed codes shared by the host table builder and the device scan kernels.
//
// The matcher the reference runs (BNDMq, /root/reference/src/pattern_matching.rs:165-209, or
// the aho-corasick DFA, src/cmd_extract.rs:260-265,332) is exact byte-string search. The device
// replaces it by "seed and verify": every text position on a grid of stride D is turned into a
// Q-base seed code; a seed that is in the query seed set is followed by an exact byte compare.
// The only property the seed code needs is
//     bytes that compare equal under the matcher  =>  equal 2-bit class,
// which holds for  cls(c) = (c >> 1) & 3  both for exact compare and for ASCII case folding
// (case is bit 5). A,C,G,T / a,c,g,t get the four distinct classes 0,1,3,2; every other byte
// (N, IUPAC, amino acids ...) aliases onto one of them and is sorted out by the verify step, so
// "N breaks the window" exactly as it does for the reference's byte compare.
//
// Texts are cut into UNITS of 16 bases (16 ASCII bytes, or 8 BAM bytes = 16 nibbles). Each unit
// packs to one 32-bit word. Two packings exist:
//   *_ord : base i of the unit sits at bits [31-2i, 30-2i] (base 0 most significant), so a seed
//           starting at any offset o inside the unit is a funnel shift of (unit, next unit);
//   *_perm: any fixed bit permutation, cheaper to compute; used when D == 16 (seed == unit).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MK_HD __host__ __device__ __forceinline__
#else
#define MK_HD inline
#endif

#define MK_UNIT_BASES 16

// ---------------------------------------------------------------------------------------------
// ASCII (1 byte / base). w0..w3 are the four little-endian 32-bit words of the 16-byte unit.
// ---------------------------------------------------------------------------------------------

// Permuted packing: 4 LOP + 3 IMAD + 1 SHF.  Bits of the result, per byte lane b (0..3) of the
// word: [8b+0,8b+1] = cls(w0.byte b), [8b+2,8b+3] = cls(w1.byte b), [8b+4,8b+5] = cls(w2.byte b),
// [8b+6,8b+7] = cls(w3.byte b).
MK_HD uint32_t mk_pack_ascii_perm(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint32_t K = 0x06060606u;
    uint32_t a = w0 & K, b = w1 & K, c = w2 & K, e = w3 & K;
    uint32_t ab = b * 4u + a;        // bits 1..4 of every byte
    uint32_t ce = e * 4u + c;        // bits 1..4 of every byte
    return ce * 8u + (ab >> 1);      // ab -> bits 0..3, ce -> bits 4..7
}

// Ordered packing: base i -> bits [31-2i, 30-2i].
MK_HD uint32_t mk_pack4_ascii_top(uint32_t w) {
    // cls of the 4 bytes gathered into the top byte, byte 0 (first base) most significant
    return (((w >> 1) & 0x03030303u) * 0x40100401u) & 0xFF000000u;
}
MK_HD uint32_t mk_pack_ascii_ord(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    return mk_pack4_ascii_top(w0) | (mk_pack4_ascii_top(w1) >> 8) | (mk_pack4_ascii_top(w2) >> 16) |
           (mk_pack4_ascii_top(w3) >> 24);
}

// ---------------------------------------------------------------------------------------------
// BAM4 (2 bases / byte, first base in the high nibble; codes "=ACMGRSVTWYHKDBN",
// /root/reference/src/cmd_tag.rs:395 decodes them through the bam crate). A unit is 8 bytes =
// two little-endian words w0, w1. Nibble class: cls4(n) = ((n2|n3) << 1) | (n1|n3), which is
// 0,1,2,3 for A(1),C(2),G(4),T(8). The host encodes query bytes to nibbles first, so text and
// query always go through the same function.
// ---------------------------------------------------------------------------------------------

// per nibble: class in bits 1..2 of the nibble
MK_HD uint32_t mk_cls4_mid(uint32_t w) {
    uint32_t lo = (w | (w >> 2)) & 0x22222222u;   // bit1 = n1|n3
    uint32_t hi = (w | (w >> 1)) & 0x44444444u;   // bit2 = n2|n3
    return lo | hi;
}
// Stride-16 seed key of a BAM unit: BAM text and the nibble-encoded queries compare by plain equality
// (no case folding: decoded BAM text is upper-case only and queries are upper-cased before encoding
// under -I), so any hash of the 8 raw bytes is a valid key — two multiply-adds instead of a class
// packing. Collisions only cost an extra exact verification.
MK_HD uint32_t mk_pack_bam_perm(uint32_t w0, uint32_t w1) {
    return w0 * 0x9E3779B1u + w1 * 0x85EBCA77u;
}

MK_HD uint32_t mk_bswap32(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0, 0x0123);
#else
    return (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
#endif
}
// 8 bases of one word -> 16 bits, first base most significant
MK_HD uint32_t mk_pack8_bam_ord(uint32_t w) {
    uint32_t c = mk_cls4_mid(mk_bswap32(w)) >> 1;       // class in bits 0..1 of every nibble
    c = (c | (c >> 2)) & 0x0F0F0F0Fu;
    c = (c | (c >> 4)) & 0x00FF00FFu;
    c = (c | (c >> 8)) & 0x0000FFFFu;
    return c;
}
MK_HD uint32_t mk_pack_bam_ord(uint32_t w0, uint32_t w1) {
    return (mk_pack8_bam_ord(w0) << 16) | mk_pack8_bam_ord(w1);
}

// ---------------------------------------------------------------------------------------------
// Seed extraction from ordered unit codes and the hashes of a seed code.
// ---------------------------------------------------------------------------------------------

// Seed of q bases starting at base offset o (0..15) of unit `cur`, `nxt` being the following unit.
MK_HD uint32_t mk_seed_ord(uint32_t cur, uint32_t nxt, uint32_t o, uint32_t q) {
#if defined(__CUDA_ARCH__)
    uint32_t win = __funnelshift_l(nxt, cur, 2u * o);
#else
    uint32_t win = o ? ((cur << (2u * o)) | (nxt >> (32u - 2u * o))) : cur;
#endif
    return win >> (32u - 2u * q);
}
// The 16-base window itself (its top 2q bits are the q-base seed)
MK_HD uint32_t mk_win_ord(uint32_t cur, uint32_t nxt, uint32_t o) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(nxt, cur, 2u * o);
#else
    return o ? ((cur << (2u * o)) | (nxt >> (32u - 2u * o))) : cur;
#endif
}

// First-level filter. Shared-memory flavour: a blocked Bloom filter of `nblocks` 64-bit blocks (any
// count; 16384 blocks = 128 KiB by default, more for large seed sets). A seed selects its block with
// mulhi(code * MK_BLOOM_MUL, nblocks) and four bit positions from a second product: two in the low
// word and two in the high word of the block — one LDS.64 per probe. Global flavour (seed sets too
// large for shared memory): a plain bitmap of 2^log2_bits bits indexed by mk_hash_f1.
#define MK_BLOOM_MUL 0x9E3779B1u
#define MK_BLOOM_MUL2 0x85EBCA77u
#define MK_BLOOM_MIN_BLOCKS 16384u
#define MK_BLOOM_MAX_BLOCKS 20480u   /* 160 KiB: beyond that the L1 left over cannot hold the loads in flight (measured) */
MK_HD uint32_t mk_bloom_block(uint32_t code, uint32_t nblocks) {
    return (uint32_t)(((uint64_t)(code * MK_BLOOM_MUL) * nblocks) >> 32);
}
// bit masks inside the low (x) and high (y) word of the block
MK_HD void mk_bloom_masks(uint32_t code, uint32_t* lo, uint32_t* hi) {
    uint32_t g = code * MK_BLOOM_MUL * MK_BLOOM_MUL2;
    *lo = (1u << ((g >> 12) & 31)) | (1u << ((g >> 17) & 31));
    *hi = (1u << ((g >> 22) & 31)) | (1u << (g >> 27));
}
// same bit positions from an already mixed word g
MK_HD void mk_bloom_masks_g(uint32_t g, uint32_t* lo, uint32_t* hi) {
    *lo = (1u << ((g >> 12) & 31)) | (1u << ((g >> 17) & 31));
    *hi = (1u << ((g >> 22) & 31)) | (1u << (g >> 27));
}
// Dual-key filter (L2-resident, stride < 16, large seed sets). Patterns long enough for a 16-base
// seed at every one of their d offsets are indexed by that 16-base seed ("long" group); the others by
// the q-base seed the shortest pattern allows ("short" group). Both keys of one text position select
// the SAME 64-bit block (through the q-base prefix they share) and differ in their bit positions, so
// one 8-byte L2 load answers both probes.
#define MK_DUAL_MUL_LONG 0xCC9E2D51u
MK_HD uint32_t mk_dual_block(uint32_t short_code, uint32_t nblocks) { return mk_bloom_block(short_code, nblocks); }
MK_HD uint32_t mk_dual_g_short(uint32_t short_code) { return short_code * MK_BLOOM_MUL * MK_BLOOM_MUL2; }
MK_HD uint32_t mk_dual_g_long(uint32_t code16) { return (code16 ^ (code16 >> 15)) * MK_DUAL_MUL_LONG; }
// mk_scan_dual8's flavour of the dual-key filter: 3 bits per key (two in the low word of the block, one in the high
// word) taken from the top bits of one product, so that a test is three shifts-by-register and one AND — this
// kernel is bound by its integer instructions as much as by its gathers (profiles/r2_cfg5_experiments/).
#define MK_DUAL3_MUL_SHORT 0x85EBCA77u
#define MK_DUAL3_MUL_LONG 0xCC9E2D51u
MK_HD uint32_t mk_dual3_g_short(uint32_t short_code) { return short_code * MK_DUAL3_MUL_SHORT; }
MK_HD uint32_t mk_dual3_g_long(uint32_t code16) { return code16 * MK_DUAL3_MUL_LONG; }
MK_HD void mk_dual3_masks(uint32_t g, uint32_t* lo, uint32_t* hi) {
    *lo = (1u << (g >> 27)) | (1u << ((g >> 22) & 31));
    *hi = 1u << ((g >> 17) & 31);
}
// key of the cuckoo seed table: the group is folded into the hashed value
MK_HD uint32_t mk_group_key(uint32_t code, uint32_t group) { return group ? (code ^ 0x3C6EF372u) : code; }
// Shared-memory flavour with 32-bit blocks and 3 bits per key (small seed sets): one LDS.32 per probe.
// `nblocks64` counts 64-bit units like the other flavour; the 32-bit block count is twice that.
MK_HD uint32_t mk_bloom32_block(uint32_t code, uint32_t nblocks64) {
    return (uint32_t)(((uint64_t)(code * MK_BLOOM_MUL) * (2u * nblocks64)) >> 32);
}
MK_HD uint32_t mk_bloom32_mask(uint32_t code) {
    uint32_t g = code * MK_BLOOM_MUL * MK_BLOOM_MUL2;
    return (1u << ((g >> 17) & 31)) | (1u << ((g >> 22) & 31)) | (1u << (g >> 27));
}
MK_HD uint32_t mk_hash_f1(uint32_t code, uint32_t log2_bits) { return (code * 0x9E3779B1u) >> (32u - log2_bits); }
// second-level filter (L2-resident bitmap probed by candidates only): bit index
MK_HD uint32_t mk_hash_f2(uint32_t code, uint32_t log2_bits) { return ((code ^ (code >> 15)) * 0x85EBCA77u) >> (32u - log2_bits); }
MK_HD uint32_t mk_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
MK_HD uint32_t mk_hash_b1(uint32_t code, uint32_t mask) { return mk_mix32(code) & mask; }
MK_HD uint32_t mk_hash_b2(uint32_t code, uint32_t mask) { return mk_mix32(code ^ 0xA5A5F00Du) & mask; }

// ASCII byte -> BAM nibble for query bytes (0xFF: the byte can never occur in decoded BAM text)
MK_HD uint8_t mk_ascii_to_nibble(uint8_t c) {
    switch (c) {
        case '=': return 0;  case 'A': return 1;  case 'C': return 2;  case 'M': return 3;
        case 'G': return 4;  case 'R': return 5;  case 'S': return 6;  case 'V': return 7;
        case 'T': return 8;  case 'W': return 9;  case 'Y': return 10; case 'H': return 11;
        case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
        default: return 0xFF;
    }
}
MK_HD uint8_t mk_fold(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c | 0x20) : c; }
