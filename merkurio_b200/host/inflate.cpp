#include "inflate.h"

#include <zlib.h>

#include <algorithm>
#include <cstddef>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mkh {

namespace {

// Table entries (both tables):
//   bits  0..7   bits to consume: the code length (first level), what is left of it (second level), or the width
//                of the first level (pointer entries); 0 = no such code
//   bits  8..11  number of extra bits (length / distance entries) or width of the second-level table (pointers)
//   bits 12..15  kLit / kEob / kPtr / kLit2
//   bits 16..31  literal, base length, base distance, or start of the second-level table
// Two literals in one entry (kLit | kLit2; first-level literal / length table only): where the code of a literal
// leaves enough of the 11 index bits to determine the symbol after it and that is a literal too, the entry holds both
// (bits 16..23 and 24..31), consumes both codes, and keeps the length of the first code in bits 8..11. Sequencing
// data is mostly literals with codes of 2-4 bits, and the table walk (load, shift, next load) is the serial chain
// that bounds the decoder: two symbols per step nearly halve it.
constexpr uint32_t kLit = 0x1000, kEob = 0x2000, kPtr = 0x4000, kLit2 = 0x8000;

constexpr uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
constexpr uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
constexpr uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
constexpr uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
constexpr uint8_t kPrecodeOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint64_t load64(const uint8_t* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64, aarch64)
}

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// kind 0: literal / length alphabet, 1: distance alphabet, 2: code-length alphabet
inline uint32_t make_entry(int kind, int sym, int nbits) {
    if (kind == 0) {
        if (sym < 256) return ((uint32_t)sym << 16) | kLit | (uint32_t)nbits;
        if (sym == 256) return kEob | (uint32_t)nbits;
        if (sym > 285) return 0;
        return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8) | (uint32_t)nbits;
    }
    if (kind == 1) {
        if (sym > 29) return 0;
        return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8) | (uint32_t)nbits;
    }
    return ((uint32_t)sym << 16) | (uint32_t)nbits;
}

// Copy a match of `len` bytes from `dist` bytes back; may write up to 7 bytes past out + len (the caller has room).
inline void copy_match_wide(uint8_t* out, size_t dist, size_t len) {
    const uint8_t* src = out - dist;
    uint8_t* const end = out + len;
    if (dist >= 8) {
        do {
            std::memcpy(out, src, 8);
            out += 8;
            src += 8;
        } while (out < end);
    } else if (dist == 1) {
        std::memset(out, *src, len);
    } else {
        // period below 8: byte by byte until a whole number of periods >= 8 lies behind, then wide
        uint8_t* o = out;
        while (o < end && (size_t)(o - out) < 8) { *o = *(o - dist); ++o; }
        const size_t step = (8 / dist) * dist;  // a multiple of the period: copying from `step` back repeats the pattern
        while (o < end) {
            std::memcpy(o, o - step, 8);  // source and destination may overlap by less than 8 only if step < 8 ...
            o += step;                    // ... so advance by `step`, not by 8: every byte written is final
        }
    }
}

}  // namespace

void Inflater::reset() {
    bitbuf_ = 0;
    bitcnt_ = 0;
    state_ = kBlockHeader;
    last_block_ = false;
    stored_left_ = 0;
    fixed_loaded_ = false;
}

// Canonical Huffman code -> two-level decode table. Returns false for over-subscribed or (other than a single
// one-bit codeword, RFC 1951 section 3.2.7) incomplete codes, and if the table would not fit.
bool Inflater::build_table(uint32_t* table, int table_bits, int table_cap, const uint8_t* lens, int n_syms, int kind) {
    int count[16] = {0};
    for (int s = 0; s < n_syms; ++s) ++count[lens[s]];
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) --max_len;
    std::memset(table, 0, sizeof(uint32_t) << table_bits);
    if (max_len == 0) return kind == 1;  // no code at all: fine for distances (a block of literals only)
    int left = 1;
    for (int len = 1; len <= 15; ++len) {
        left = (left << 1) - count[len];
        if (left < 0) return false;
    }
    if (left > 0 && max_len != 1) return false;  // (zlib accepts the same: a code that consists of one 1-bit codeword)

    int offs[17];
    offs[1] = 0;
    for (int len = 1; len <= 15; ++len) offs[len + 1] = offs[len] + count[len];
    uint16_t sorted[288];
    {
        int pos[16];
        for (int len = 1; len <= 15; ++len) pos[len] = offs[len];
        for (int s = 0; s < n_syms; ++s)
            if (lens[s]) sorted[pos[lens[s]]++] = (uint16_t)s;
    }

    const uint32_t first_mask = (1u << table_bits) - 1;
    uint32_t code = 0;  // canonical code of the current symbol, most significant bit first
    int next_free = 1 << table_bits;
    uint32_t cur_prefix = ~0u;
    int sub_start = 0, sub_bits = 0;
    for (int len = 1; len <= max_len; ++len) {
        for (int k = 0; k < count[len]; ++k, ++code) {
            const int sym = sorted[offs[len] + k];
            const uint32_t rev = reverse_bits(code, len);
            if (len <= table_bits) {
                const uint32_t e = make_entry(kind, sym, len);
                for (uint32_t i = rev; i <= first_mask; i += 1u << len) table[i] = e;
                continue;
            }
            const uint32_t prefix = rev & first_mask;
            if (prefix != cur_prefix) {
                // a new second-level table: wide enough for the longest code that shares this prefix
                cur_prefix = prefix;
                sub_bits = len - table_bits;
                int room = 1 << sub_bits, l = len, n_here = count[len] - k;
                while (sub_bits + table_bits < max_len) {
                    room -= n_here;
                    if (room <= 0) break;
                    ++sub_bits;
                    ++l;
                    room <<= 1;
                    n_here = count[l];
                }
                sub_start = next_free;
                next_free += 1 << sub_bits;
                if (next_free > table_cap) return false;
                std::memset(table + sub_start, 0, sizeof(uint32_t) << sub_bits);
                table[prefix] = ((uint32_t)sub_start << 16) | ((uint32_t)sub_bits << 8) | kPtr | (uint32_t)table_bits;
            }
            const uint32_t e = make_entry(kind, sym, len - table_bits);
            for (uint32_t i = rev >> table_bits; i < (1u << sub_bits); i += 1u << (len - table_bits)) table[sub_start + i] = e;
        }
        code <<= 1;
    }
    if (kind == 0) {
        // pair up literals (ascending index: the entry of the remaining bits, a smaller index, may be a pair already —
        // its first literal and first code length are what is needed)
        for (uint32_t i = 0; i <= first_mask; ++i) {
            const uint32_t e = table[i];
            if ((e & (kLit | kPtr)) != kLit) continue;
            const uint32_t l1 = e & 0xFF;
            if ((int)l1 >= table_bits) continue;
            const uint32_t e2 = table[i >> l1];
            if ((e2 & (kLit | kPtr)) != kLit) continue;
            const uint32_t l2 = (e2 & kLit2) ? ((e2 >> 8) & 15) : (e2 & 0xFF);
            if (l1 + l2 > (uint32_t)table_bits) continue;
            table[i] = ((((e >> 16) & 0xFF) | (((e2 >> 16) & 0xFF) << 8)) << 16) | kLit | kLit2 | (l1 << 8) | (l1 + l2);
        }
    }
    return true;
}

void Inflater::use_fixed_codes() {
    if (fixed_loaded_) return;
    uint8_t lens[288 + 32];
    int s = 0;
    for (; s < 144; ++s) lens[s] = 8;
    for (; s < 256; ++s) lens[s] = 9;
    for (; s < 280; ++s) lens[s] = 7;
    for (; s < 288; ++s) lens[s] = 8;
    build_table(litlen_, kLitlenBits, kLitlenEntries, lens, 288, 0);
    for (s = 0; s < 32; ++s) lens[s] = 5;
    build_table(dist_, kDistBits, kDistEntries, lens, 32, 1);
    fixed_loaded_ = true;
}

#define MK_REFILL()                                  \
    do {                                             \
        bitbuf |= load64(in) << bitcnt;              \
        in += (63 - bitcnt) >> 3;                    \
        bitcnt |= 56;                                \
    } while (0)
#define MK_BITS(n) ((uint32_t)bitbuf & ((1u << (n)) - 1u))
#define MK_DROP(n) (bitbuf >>= (n), bitcnt -= (n))

// The header of a dynamic block (the caller has made sure that it is in the buffer in full, or that the input ends).
bool Inflater::read_dynamic_header(const uint8_t*& in_ref, const uint8_t* in_end, bool in_final) {
    const uint8_t* in = in_ref;
    uint64_t bitbuf = bitbuf_;
    unsigned bitcnt = bitcnt_;
    MK_REFILL();
    const unsigned hlit = MK_BITS(5) + 257; MK_DROP(5);
    const unsigned hdist = MK_BITS(5) + 1; MK_DROP(5);
    const unsigned hclen = MK_BITS(4) + 4; MK_DROP(4);
    if (hlit > 286 || hdist > 30) return false;
    uint8_t pre_lens[19] = {0};
    for (unsigned i = 0; i < hclen; ++i) {
        if (bitcnt < 3) MK_REFILL();
        pre_lens[kPrecodeOrder[i]] = (uint8_t)MK_BITS(3);
        MK_DROP(3);
    }
    uint32_t pre[128];
    if (!build_table(pre, 7, 128, pre_lens, 19, 2)) return false;
    uint8_t lens[286 + 30 + 138];
    unsigned n = 0;
    const unsigned total = hlit + hdist;
    while (n < total) {
        MK_REFILL();
        const uint32_t e = pre[bitbuf & 127];
        const unsigned nb = e & 0xFF;
        if (!nb) return false;
        MK_DROP(nb);
        const unsigned sym = e >> 16;
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
        } else if (sym == 16) {
            if (n == 0) return false;
            const unsigned rep = 3 + MK_BITS(2); MK_DROP(2);
            std::memset(lens + n, lens[n - 1], rep);
            n += rep;
        } else if (sym == 17) {
            const unsigned rep = 3 + MK_BITS(3); MK_DROP(3);
            std::memset(lens + n, 0, rep);
            n += rep;
        } else {
            const unsigned rep = 11 + MK_BITS(7); MK_DROP(7);
            std::memset(lens + n, 0, rep);
            n += rep;
        }
    }
    if (n != total) return false;       // a repeat ran over the end
    if (lens[256] == 0) return false;   // no end-of-block code
    if (in_final && in - (bitcnt >> 3) > in_end) return false;  // the header runs past the end of the input
    fixed_loaded_ = false;  // (before the tables are touched: a header that fails half way must not leave them marked as the fixed ones)
    if (!build_table(litlen_, kLitlenBits, kLitlenEntries, lens, (int)hlit, 0)) return false;
    if (!build_table(dist_, kDistBits, kDistEntries, lens + hlit, (int)hdist, 1)) return false;
    in_ref = in;
    bitbuf_ = bitbuf;
    bitcnt_ = bitcnt;
    return true;
}

__attribute__((always_inline)) inline Inflater::Status Inflater::run_impl(const uint8_t** in_p, const uint8_t* in_end, bool in_final,
                                                                         const uint8_t* out_base, uint8_t** out_p, uint8_t* out_end) {
    const uint8_t* in = *in_p;
    uint8_t* out = *out_p;
    uint64_t bitbuf = bitbuf_;
    unsigned bitcnt = bitcnt_;
    Status rc = kError;
    const uint8_t* const real_end = in_end;
    bool in_tail = false;
    constexpr uint32_t kLMask = (1u << kLitlenBits) - 1, kDMask = (1u << kDistBits) - 1;

#define MK_LEAVE(status) do { rc = (status); goto leave; } while (0)

    for (;;) {
        if (state_ == kDone) MK_LEAVE(kStreamEnd);
        if (state_ == kBlockHeader) {
            if (in_end - in < (ptrdiff_t)kInputMargin) {
                if (!in_final) MK_LEAVE(kNeedInput);
                // whole bytes still in the bit buffer go back first: `in` must be the true position for the check and
                // the copy below (a stored block gives its bytes back as well, and must not step in front of the copy)
                in -= bitcnt >> 3;
                bitcnt &= 7;
                bitbuf &= (1ull << bitcnt) - 1;
                if (in > in_end) MK_LEAVE(kError);
                if (!in_tail) {
                    // the input ends within a header's length: go on in a zero-padded copy, so that nothing is read
                    // behind the caller's buffer if the stream is cut inside a block header
                    const size_t n = (size_t)(in_end - in);
                    std::memcpy(tail_, in, n);
                    std::memset(tail_ + n, 0, sizeof tail_ - n);
                    in = tail_;
                    in_end = tail_ + n;
                    in_tail = true;
                }
            }
            MK_REFILL();
            last_block_ = MK_BITS(1); MK_DROP(1);
            const unsigned type = MK_BITS(2); MK_DROP(2);
            if (type == 0) {
                // stored: drop the rest of the current byte, give back the whole bytes still in the bit buffer
                MK_DROP(bitcnt & 7);
                in -= bitcnt >> 3;
                bitbuf = 0;
                bitcnt = 0;
                if (in_end - in < 4) MK_LEAVE(kError);  // (in_final, or the margin above would have held)
                const uint32_t len = in[0] | ((uint32_t)in[1] << 8), nlen = in[2] | ((uint32_t)in[3] << 8);
                if ((len ^ nlen) != 0xFFFFu) MK_LEAVE(kError);
                in += 4;
                stored_left_ = len;
                state_ = kStored;
            } else if (type == 1) {
                use_fixed_codes();
                state_ = kHuffman;
            } else if (type == 2) {
                bitbuf_ = bitbuf;
                bitcnt_ = bitcnt;
                if (!read_dynamic_header(in, in_end, in_final)) MK_LEAVE(kError);
                bitbuf = bitbuf_;
                bitcnt = bitcnt_;
                state_ = kHuffman;
            } else {
                MK_LEAVE(kError);
            }
        }
        if (state_ == kStored) {
            while (stored_left_) {
                if (in >= in_end) MK_LEAVE(in_final ? kError : kNeedInput);
                if (out >= out_end) MK_LEAVE(kOutputFull);
                size_t n = stored_left_;
                if (n > (size_t)(in_end - in)) n = (size_t)(in_end - in);
                if (n > (size_t)(out_end - out)) n = (size_t)(out_end - out);
                std::memcpy(out, in, n);
                in += n;
                out += n;
                stored_left_ -= (uint32_t)n;
            }
            state_ = last_block_ ? kDone : kBlockHeader;
            if (stop_at_block_end_ && !last_block_) MK_LEAVE(kBlockEnd);
            continue;
        }
        // state_ == kHuffman
        for (;;) {
            // fast path: far from the end of both buffers, no checks per symbol. Invariants at the top of the loop: at
            // least 56 bits in the buffer, `e` is the first-level entry of the next bits.
            if (in_end - in >= 32 && out_end - out >= (ptrdiff_t)kOutputMargin) {  // (signed: `in` may be behind in_end at the end of the input)
                MK_REFILL();
                uint32_t e = litlen_[bitbuf & kLMask];
                do {
                    if (e & kLit) {
                        // up to three entries (<= 33 bits) = up to six literals; both bytes of an entry are stored
                        // whether or not it holds two (the second one is overwritten if it does not)
#define MK_EMIT_LITERALS()                                   \
    do {                                                     \
        MK_DROP(e & 0xFF);                                   \
        const uint16_t two = (uint16_t)(e >> 16);            \
        std::memcpy(out, &two, 2);                           \
        out += 1 + ((e >> 15) & 1);                          \
        e = litlen_[bitbuf & kLMask];                        \
    } while (0)
                        MK_EMIT_LITERALS();
                        if (e & kLit) {
                            MK_EMIT_LITERALS();
                            if (e & kLit) MK_EMIT_LITERALS();
                        }
#undef MK_EMIT_LITERALS
                        MK_REFILL();  // (>= 23 bits were left: the index bits of `e` stay what they are)
                        continue;
                    }
                    if (e & kPtr) {
                        MK_DROP(kLitlenBits);
                        e = litlen_[(e >> 16) + MK_BITS((e >> 8) & 15)];
                    }
                    unsigned nb = e & 0xFF;
                    if (!nb) MK_LEAVE(kError);
                    MK_DROP(nb);
                    if (e & kLit) {  // (second level: one literal)
                        *out++ = (uint8_t)(e >> 16);
                        MK_REFILL();
                        e = litlen_[bitbuf & kLMask];
                        continue;
                    }
                    if (e & kEob) goto block_done;
                    const unsigned xl = (e >> 8) & 15;
                    const size_t len = (e >> 16) + MK_BITS(xl);
                    MK_DROP(xl);
                    uint32_t d = dist_[bitbuf & kDMask];
                    if (d & kPtr) {
                        MK_DROP(kDistBits);
                        d = dist_[(d >> 16) + MK_BITS((d >> 8) & 15)];
                    }
                    nb = d & 0xFF;
                    if (!nb) MK_LEAVE(kError);
                    MK_DROP(nb);
                    const unsigned xd = (d >> 8) & 15;  // (15 + 5 + 15 bits are gone at most: 21 are left for these <= 13)
                    const size_t dist = (d >> 16) + MK_BITS(xd);
                    MK_DROP(xd);
                    if (dist > (size_t)(out - out_base)) MK_LEAVE(kError);
                    // the next entry is on its way while the match is copied
                    MK_REFILL();
                    e = litlen_[bitbuf & kLMask];
                    const uint8_t* src = out - dist;
                    if (dist >= 8) {
                        // matches in sequencing data are short (6-10 bases found again within 32 KiB): 16 bytes without a loop
                        std::memcpy(out, src, 8);
                        std::memcpy(out + 8, src + 8, 8);
                        if (len > 16) {
                            size_t k = 16;
                            do {
                                std::memcpy(out + k, src + k, 8);
                                std::memcpy(out + k + 8, src + k + 8, 8);
                                k += 16;
                            } while (k < len);
                        }
                    } else {
                        copy_match_wide(out, dist, len);
                    }
                    out += len;
                } while (in_end - in >= 32 && out_end - out >= (ptrdiff_t)kOutputMargin);
                // (`e` is looked up again below: none of its bits were consumed)
            }
            // careful path: one symbol, with every check
            if (!in_final && in_end - in < 32) MK_LEAVE(kNeedInput);
            if (out_end - out < (ptrdiff_t)kOutputMargin && !(in_final && exact_tail_)) MK_LEAVE(kOutputFull);
            if (in_final && in - (bitcnt >> 3) > in_end) MK_LEAVE(kError);  // decoding zeros behind a truncated stream
            MK_REFILL();
            uint32_t e = litlen_[bitbuf & kLMask];
            if (e & kPtr) {
                MK_DROP(kLitlenBits);
                e = litlen_[(e >> 16) + MK_BITS((e >> 8) & 15)];
            }
            unsigned nb = e & 0xFF;
            if (!nb) MK_LEAVE(kError);
            if (e & kLit2) nb = (e >> 8) & 15;  // here one literal at a time: the first code of the pair only
            MK_DROP(nb);
            if (e & kLit) {
                if (in_final && in - (bitcnt >> 3) > in_end) MK_LEAVE(kError);  // (decoded from behind the end: not output)
                if (out >= out_end) MK_LEAVE(kError);
                *out++ = (uint8_t)(e >> 16);
                continue;
            }
            if (e & kEob) goto block_done;
            const unsigned xl = (e >> 8) & 15;
            const size_t len = (e >> 16) + MK_BITS(xl);
            MK_DROP(xl);
            uint32_t d = dist_[bitbuf & kDMask];
            if (d & kPtr) {
                MK_DROP(kDistBits);
                d = dist_[(d >> 16) + MK_BITS((d >> 8) & 15)];
            }
            nb = d & 0xFF;
            if (!nb) MK_LEAVE(kError);
            MK_DROP(nb);
            const unsigned xd = (d >> 8) & 15;
            if (bitcnt < xd) MK_REFILL();
            const size_t dist = (d >> 16) + MK_BITS(xd);
            MK_DROP(xd);
            if (in_final && in - (bitcnt >> 3) > in_end) MK_LEAVE(kError);
            if (dist > (size_t)(out - out_base)) MK_LEAVE(kError);
            if (len > (size_t)(out_end - out)) MK_LEAVE(kError);  // (exact tail only: more output than the caller expects)
            for (size_t i = 0; i < len; ++i) out[i] = out[i - dist];
            out += len;
        }
    block_done:
        if (in_final && in - (bitcnt >> 3) > in_end) MK_LEAVE(kError);
        if (last_block_) {
            // the stream ends with the current byte: hand the whole bytes still in the bit buffer back
            MK_DROP(bitcnt & 7);
            in -= bitcnt >> 3;
            bitbuf = 0;
            bitcnt = 0;
            state_ = kDone;
        } else {
            state_ = kBlockHeader;
            if (stop_at_block_end_) MK_LEAVE(kBlockEnd);
        }
    }

leave:
    // whole bytes still in the bit buffer go back to the input: at most 7 bits stay behind between calls
    if (rc != kError) {
        in -= bitcnt >> 3;
        bitcnt &= 7;
        bitbuf &= (1ull << bitcnt) - 1;
    }
    if (in_tail) in = real_end - (in_end - in);
    *in_p = in;
    *out_p = out;
    bitbuf_ = bitbuf;
    bitcnt_ = bitcnt;
    return rc;
#undef MK_LEAVE
}

#undef MK_REFILL
#undef MK_BITS
#undef MK_DROP

// The decoder proper is compiled twice: as it is, and for CPUs with BMI2 (shifts by a register without the detour
// through %cl, bzhi for the masks: 1.3x on sequencing data), picked at run time.
Inflater::Status Inflater::run(const uint8_t** in_p, const uint8_t* in_end, bool in_final, const uint8_t* out_base, uint8_t** out_p,
                               uint8_t* out_end) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool bmi2 = __builtin_cpu_supports("bmi2");
    if (bmi2) return run_bmi2(in_p, in_end, in_final, out_base, out_p, out_end);
#endif
    return run_generic(in_p, in_end, in_final, out_base, out_p, out_end);
}
Inflater::Status Inflater::run_generic(const uint8_t** in_p, const uint8_t* in_end, bool in_final, const uint8_t* out_base,
                                       uint8_t** out_p, uint8_t* out_end) {
    return run_impl(in_p, in_end, in_final, out_base, out_p, out_end);
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2")))
#endif
Inflater::Status Inflater::run_bmi2(const uint8_t** in_p, const uint8_t* in_end, bool in_final, const uint8_t* out_base,
                                    uint8_t** out_p, uint8_t* out_end) {
    return run_impl(in_p, in_end, in_final, out_base, out_p, out_end);
}

// ---------------------------------------------------------------------------------------------------------------
// A stream taken apart at block boundaries
// ---------------------------------------------------------------------------------------------------------------
#define MK_REFILL()                                  \
    do {                                             \
        bitbuf |= load64(in) << bitcnt;              \
        in += (63 - bitcnt) >> 3;                    \
        bitcnt |= 56;                                \
    } while (0)
#define MK_BITS(n) ((uint32_t)bitbuf & ((1u << (n)) - 1u))
#define MK_DROP(n) (bitbuf >>= (n), bitcnt -= (n))

const uint8_t* Inflater::seek_bits(const uint8_t* base, uint64_t bitpos) {
    const uint8_t* in = base + (bitpos >> 3);
    const unsigned sub = (unsigned)(bitpos & 7);
    bitbuf_ = 0;
    bitcnt_ = 0;
    if (sub) {
        bitbuf_ = (uint64_t)(*in >> sub);
        bitcnt_ = 8 - sub;
        ++in;
    }
    state_ = kBlockHeader;
    last_block_ = false;
    stored_left_ = 0;
    return in;
}

bool Inflater::probe_dynamic_header(const uint8_t* base, const uint8_t* end, uint64_t bitpos) {
    const uint8_t* p = base + (bitpos >> 3);
    if (end - p < 8) return false;
    const unsigned sub = (unsigned)(bitpos & 7);
    const uint64_t v = load64(p) >> sub;  // >= 57 bits
    // BFINAL = 0, BTYPE = 2 (bits 1..2, least significant first), HLIT <= 29, HDIST <= 29
    if ((v & 7) != 4 || ((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) return false;
    // the code-length code must be complete (Kraft sum 1): what zlib and every other deflate writes
    const unsigned hclen = (unsigned)((v >> 13) & 15) + 4;
    uint64_t w = v >> 17;  // 40 bits = 13 lengths; the rest from a second load
    unsigned kraft = 0;
    for (unsigned i = 0; i < hclen; ++i) {
        if (i == 13) w = load64(base + ((bitpos + 17 + 39) >> 3)) >> ((bitpos + 17 + 39) & 7);
        const unsigned len = (unsigned)(w & 7);
        w >>= 3;
        if (len) kraft += 128u >> len;
    }
    if (kraft != 128) return false;
    // the whole header: both codes built (complete, not over-subscribed, end-of-block code present)
    const uint8_t* in = seek_bits(base, bitpos);
    uint64_t bitbuf = bitbuf_;
    unsigned bitcnt = bitcnt_;
    MK_REFILL();
    MK_DROP(3);
    bitbuf_ = bitbuf;
    bitcnt_ = bitcnt;
    return read_dynamic_header(in, end, true);
}

Inflater::MarkerRun Inflater::run_markers(const uint8_t* base, const uint8_t* end, uint64_t start_bit, uint64_t stop_bit,
                                          std::vector<uint16_t>* outv, size_t max_symbols) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool bmi2 = __builtin_cpu_supports("bmi2");
    if (bmi2) return run_markers_bmi2(base, end, start_bit, stop_bit, outv, max_symbols);
#endif
    return run_markers_impl(base, end, start_bit, stop_bit, outv, max_symbols);
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2")))
#endif
Inflater::MarkerRun Inflater::run_markers_bmi2(const uint8_t* base, const uint8_t* end, uint64_t start_bit, uint64_t stop_bit,
                                               std::vector<uint16_t>* outv, size_t max_symbols) {
    return run_markers_impl(base, end, start_bit, stop_bit, outv, max_symbols);
}
__attribute__((always_inline)) inline Inflater::MarkerRun Inflater::run_markers_impl(const uint8_t* base, const uint8_t* end, uint64_t start_bit,
                                                                                uint64_t stop_bit, std::vector<uint16_t>* outv, size_t max_symbols) {
    MarkerRun res;
    constexpr size_t kWin = 32768;
    constexpr uint32_t kLMask = (1u << kLitlenBits) - 1, kDMask = (1u << kDistBits) - 1;
    std::vector<uint16_t>& out = *outv;
    if (out.size() < kWin + (1u << 20)) out.resize(kWin + (1u << 20));
    for (size_t j = 0; j < kWin; ++j) out[j] = (uint16_t)(256 + j);
    size_t n = kWin;
    const uint8_t* in = seek_bits(base, start_bit);
    uint64_t bitbuf = bitbuf_;
    unsigned bitcnt = bitcnt_;
    for (;;) {
        if (in - (bitcnt >> 3) > end) return res;
        MK_REFILL();
        const bool last = MK_BITS(1); MK_DROP(1);
        const unsigned type = MK_BITS(2); MK_DROP(2);
        if (type == 0) {
            MK_DROP(bitcnt & 7);
            in -= bitcnt >> 3;
            bitbuf = 0;
            bitcnt = 0;
            if (end - in < 4) return res;
            const uint32_t len = in[0] | ((uint32_t)in[1] << 8), nlen = in[2] | ((uint32_t)in[3] << 8);
            if ((len ^ nlen) != 0xFFFFu) return res;
            in += 4;
            if ((size_t)(end - in) < len) return res;
            if (n - kWin + len > max_symbols) return res;
            if (out.size() < n + len) out.resize(std::max(out.size() * 2, n + len));
            for (uint32_t i = 0; i < len; ++i) out[n + i] = in[i];
            n += len;
            in += len;
        } else if (type == 3) {
            return res;
        } else {
            if (type == 1) {
                use_fixed_codes();
            } else {
                bitbuf_ = bitbuf;
                bitcnt_ = bitcnt;
                if (!read_dynamic_header(in, end, true)) return res;
                bitbuf = bitbuf_;
                bitcnt = bitcnt_;
            }
            bool block_done = false;
            while (!block_done) {
                if (in - (bitcnt >> 3) > end) return res;
                if (out.size() < n + 4096) {
                    if (n - kWin > max_symbols) return res;
                    out.resize(out.size() * 2);
                }
                // as long as 32 input bytes and 600 output symbols are at hand: no checks per symbol; the entry of the
                // next symbol is looked up before the current match is copied
                uint16_t* o = out.data() + n;
                uint16_t* const o_stop = out.data() + out.size() - 600;
                MK_REFILL();
                uint32_t e = litlen_[bitbuf & kLMask];
                while (end - in >= 32 && o < o_stop) {
                    if (e & kLit) {
                        // up to three first-level entries = six literals without another refill (as in the byte decoder)
#define MK_EMIT_LITERALS()                                   \
    do {                                                     \
        MK_DROP(e & 0xFF);                                   \
        o[0] = (uint16_t)((e >> 16) & 0xFF);                 \
        o[1] = (uint16_t)(e >> 24);                          \
        o += 1 + ((e >> 15) & 1);                            \
        e = litlen_[bitbuf & kLMask];                        \
    } while (0)
                        MK_EMIT_LITERALS();
                        if (e & kLit) {
                            MK_EMIT_LITERALS();
                            if (e & kLit) MK_EMIT_LITERALS();
                        }
#undef MK_EMIT_LITERALS
                        MK_REFILL();
                        continue;
                    }
                    if (e & kPtr) {
                        MK_DROP(kLitlenBits);
                        e = litlen_[(e >> 16) + MK_BITS((e >> 8) & 15)];
                    }
                    unsigned nb = e & 0xFF;
                    if (!nb) return res;
                    MK_DROP(nb);
                    if (e & kLit) {
                        *o++ = (uint16_t)((e >> 16) & 0xFF);
                        MK_REFILL();
                        e = litlen_[bitbuf & kLMask];
                        continue;
                    }
                    if (e & kEob) { block_done = true; break; }
                    const unsigned xl = (e >> 8) & 15;
                    const size_t len = (e >> 16) + MK_BITS(xl);
                    MK_DROP(xl);
                    uint32_t d = dist_[bitbuf & kDMask];
                    if (d & kPtr) {
                        MK_DROP(kDistBits);
                        d = dist_[(d >> 16) + MK_BITS((d >> 8) & 15)];
                    }
                    nb = d & 0xFF;
                    if (!nb) return res;
                    MK_DROP(nb);
                    const unsigned xd = (d >> 8) & 15;
                    const size_t dist = (d >> 16) + MK_BITS(xd);
                    MK_DROP(xd);
                    MK_REFILL();
                    e = litlen_[bitbuf & kLMask];
                    // (dist <= 32768 <= o - out.data() always: in front of the start lies the window of place holders)
                    const uint16_t* src = o - dist;
                    if (dist >= 8) {
                        std::memcpy(o, src, 16);  // eight symbols at a time; 600 symbols of room behind o
                        if (len > 8) {
                            size_t k = 8;
                            do {
                                std::memcpy(o + k, src + k, 16);
                                k += 8;
                            } while (k < len);
                        }
                    } else {
                        for (size_t k = 0; k < len; ++k) o[k] = src[k];
                    }
                    o += len;
                }
                n = (size_t)(o - out.data());
                if (block_done) break;
                if (end - in >= 32) continue;  // the output buffer was the reason: grow it
                // the last bytes of the input: one symbol at a time, with every check (`e` is looked up again: none of
                // its bits were consumed)
                for (;;) {
                    if (in - (bitcnt >> 3) > end) return res;
                    if (out.size() < n + 600) break;
                    MK_REFILL();
                    e = litlen_[bitbuf & kLMask];
                    if (e & kPtr) {
                        MK_DROP(kLitlenBits);
                        e = litlen_[(e >> 16) + MK_BITS((e >> 8) & 15)];
                    }
                    unsigned nb = e & 0xFF;
                    if (!nb) return res;
                    if (e & kLit2) nb = (e >> 8) & 15;  // one literal at a time here
                    MK_DROP(nb);
                    if (e & kLit) {
                        out[n++] = (uint16_t)((e >> 16) & 0xFF);
                        continue;
                    }
                    if (e & kEob) { block_done = true; break; }
                    const unsigned xl = (e >> 8) & 15;
                    const size_t len = (e >> 16) + MK_BITS(xl);
                    MK_DROP(xl);
                    uint32_t d = dist_[bitbuf & kDMask];
                    if (d & kPtr) {
                        MK_DROP(kDistBits);
                        d = dist_[(d >> 16) + MK_BITS((d >> 8) & 15)];
                    }
                    nb = d & 0xFF;
                    if (!nb) return res;
                    MK_DROP(nb);
                    const unsigned xd = (d >> 8) & 15;
                    if (bitcnt < xd) MK_REFILL();
                    const size_t dist = (d >> 16) + MK_BITS(xd);
                    MK_DROP(xd);
                    for (size_t k = 0; k < len; ++k) out[n + k] = out[n + k - dist];
                    n += len;
                }
            }
        }
        if (in - (bitcnt >> 3) > end) return res;
        const uint64_t pos = (uint64_t)(in - base) * 8 - bitcnt;
        if (last || pos >= stop_bit) {
            res.n_out = n;
            res.ok = true;
            res.ended_final = last;
            res.end_bit = pos;
            return res;
        }
    }
}

#undef MK_REFILL
#undef MK_BITS
#undef MK_DROP

int parse_gzip_header(const uint8_t* p, size_t n, size_t* len) {
    if (n < 10) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return -1;
    const unsigned flg = p[3];
    size_t q = 10;
    if (flg & 4) {  // FEXTRA
        if (n < q + 2) return 0;
        const size_t xlen = p[q] | ((size_t)p[q + 1] << 8);
        q += 2;
        if (n < q + xlen) return 0;
        q += xlen;
    }
    for (unsigned bit : {8u, 16u}) {  // FNAME, FCOMMENT: zero-terminated
        if (!(flg & bit)) continue;
        const void* z = std::memchr(p + q, 0, n - q);
        if (!z) return 0;
        q = (size_t)(static_cast<const uint8_t*>(z) - p) + 1;
    }
    if (flg & 2) {  // FHCRC: the low 16 bits of the CRC-32 of the header so far (checked, as zlib and flate2 do)
        if (n < q + 2) return 0;
        const uint32_t want = p[q] | ((uint32_t)p[q + 1] << 8);
        if ((crc32_fast(0, p, q) & 0xFFFFu) != want) return -1;
        q += 2;
    }
    *len = q;
    return 1;
}

// Symbols -> bytes through one table: entries 0..255 are the bytes themselves, entry 256 + j is byte j of the window, so a
// symbol is one load whatever it is (no branch on "is it a place holder": they make up a tenth of sequencing data,
// spread evenly).
#if defined(__x86_64__) && defined(__GNUC__)
// sixteen symbols per step where none of them is a place holder
__attribute__((target("avx2")))
static size_t resolve_markers_avx2(const uint16_t* src, size_t n, const uint8_t* lut, uint32_t first_valid_symbol, uint8_t* dst, uint32_t* bad) {
    size_t i = 0;
    const __m256i high = _mm256_set1_epi16((short)0xFF00);
    uint32_t b = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        if (_mm256_testz_si256(v, high)) {
            const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(v, v), 0x08);  // lanes 0 and 2 hold the 16 bytes
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm256_castsi256_si128(p));
            continue;
        }
        for (int k = 0; k < 16; ++k) {
            const uint32_t x = src[i + k];
            dst[i + k] = lut[x];
            b |= (uint32_t)(x >= 256) & (uint32_t)(x < first_valid_symbol);
        }
    }
    *bad |= b;
    return i;
}
#endif

bool resolve_markers(const uint16_t* src, size_t n, const uint8_t* window, size_t window_valid, uint8_t* dst) {
    static thread_local uint8_t lut[256 + 32768];
    static thread_local bool lut_ready = false;
    if (!lut_ready) {
        for (int v = 0; v < 256; ++v) lut[v] = (uint8_t)v;
        lut_ready = true;
    }
    std::memcpy(lut + 256, window, 32768);
    const uint32_t first_valid_symbol = 256 + (uint32_t)(32768 - window_valid);  // place holders below it point in front of the data
    uint32_t bad = 0;
    size_t i = 0;
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) i = resolve_markers_avx2(src, n, lut, first_valid_symbol, dst, &bad);
#endif
    for (; i < n; ++i) {
        const uint32_t x = src[i];
        dst[i] = lut[x];
        bad |= (uint32_t)(x >= 256) & (uint32_t)(x < first_valid_symbol);
    }
    return bad == 0;
}

#if defined(__x86_64__) && defined(__GNUC__)
// CRC-32 (IEEE 802.3, reflected) of a buffer whose length is a multiple of 16 and at least 64, by carry-less
// multiplication: four 128-bit lanes folded over 64 bytes per step, reduced to 32 bits at the end (Gopal et al.,
// "Fast CRC computation for generic polynomials using PCLMULQDQ", Intel 2009). crc is the running value as zlib's
// crc32() takes and returns it.
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(uint32_t crc, const uint8_t* p, size_t len) {
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);
    const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0)), x2 = _mm_loadu_si128((const __m128i*)(p + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32)), x4 = _mm_loadu_si128((const __m128i*)(p + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)~crc));
    p += 64; len -= 64;
    while (len >= 64) {
        __m128i t1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), t2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        __m128i t3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), t4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11); x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11); x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, t1), _mm_loadu_si128((const __m128i*)(p + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, t2), _mm_loadu_si128((const __m128i*)(p + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, t3), _mm_loadu_si128((const __m128i*)(p + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, t4), _mm_loadu_si128((const __m128i*)(p + 48)));
        p += 64; len -= 64;
    }
    // four lanes -> one
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), t), x2);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), t), x3);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), t), x4);
    while (len >= 16) {
        t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
        x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), t), _mm_loadu_si128((const __m128i*)p));
        p += 16; len -= 16;
    }
    // 128 -> 64 bits
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    x2 = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, mask32);
    x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5, 0x00), x2);
    // Barrett reduction 64 -> 32 bits
    x2 = _mm_and_si128(x1, mask32);
    x2 = _mm_clmulepi64_si128(x2, poly, 0x10);
    x2 = _mm_and_si128(x2, mask32);
    x2 = _mm_clmulepi64_si128(x2, poly, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return ~(uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t len) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (clmul && len >= 64) {
        const size_t bulk = len & ~(size_t)15;
        crc = crc32_clmul(crc, p, bulk);
        p += bulk;
        len -= bulk;
    }
#endif
    while (len) {  // (zlib takes 32-bit lengths)
        const size_t n = len < (1u << 30) ? len : (1u << 30);
        crc = (uint32_t)crc32(crc, p, (uInt)n);
        p += n;
        len -= n;
    }
    return crc;
}

bool inflate_exact(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    Inflater inf;
    inf.set_exact_tail(true);
    const uint8_t* ip = in;
    uint8_t* op = out;
    const Inflater::Status rc = inf.run(&ip, in + in_len, true, out, &op, out + out_len);
    return rc == Inflater::kStreamEnd && op == out + out_len;
}

}  // namespace mkh
