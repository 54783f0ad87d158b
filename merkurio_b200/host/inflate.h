// Raw DEFLATE (RFC 1951) decoder for the ingest path: gzip members and BGZF blocks of FASTA / FASTQ / BAM inputs
// (the reference reads them through needletail / the bam crate, i.e. flate2). zlib's inflate decodes one symbol
// per table walk through a 32-bit bit buffer; on sequencing data (mostly literals and short matches) that is the
// slowest stage of the whole pipeline once the matching runs on the GPU. This decoder keeps 56-63 bits in a
// 64-bit buffer (one unaligned load per refill), resolves literal / length codes through an 11-bit first-level
// table, and copies matches eight bytes at a time. It can stop and resume at any symbol boundary, so a stream
// is decoded in pieces of bounded size from an input buffer that is refilled in between.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mkh {

class Inflater {
public:
    enum Status {
        kNeedInput,   // fewer than kInputMargin bytes left and `in_final` is false: call again with more input
        kOutputFull,  // fewer than kOutputMargin bytes of room left: hand out what was produced and call again
        kStreamEnd,   // the final block ended; `in` points behind the last byte the stream used
        kError,       // invalid or (with in_final) truncated stream
        kBlockEnd     // (only with set_stop_at_block_end) a block that is not the last one ended; tell_bits() says where
    };
    // A call makes progress unless it returns kNeedInput / kOutputFull with margins as below.
    static constexpr size_t kInputMargin = 1200;  // a dynamic block header is parsed only when it is there in full
    static constexpr size_t kOutputMargin = 320;  // a match of 258 bytes + the overshoot of the wide copies

    Inflater() { reset(); }
    void reset();
    // With in_final: the output buffer is exactly as large as the stream's output, so decode up to its last byte
    // (byte-wise copies near the end) instead of stopping kOutputMargin bytes short of it.
    void set_exact_tail(bool on) { exact_tail_ = on; }

    // --- a stream taken apart at block boundaries (pgzip.cpp: one gzip member decoded by several threads) -----------
    void set_stop_at_block_end(bool on) { stop_at_block_end_ = on; }
    // Continue at bit `bitpos` of the buffer that starts at `base`, where a block header stands; returns the `in` to pass on.
    const uint8_t* seek_bits(const uint8_t* base, uint64_t bitpos);
    // Bit position of the next unread bit, for the `in` run() left behind.
    uint64_t tell_bits(const uint8_t* base, const uint8_t* in) const { return (uint64_t)(in - base) * 8 - bitcnt_; }
    // Is there the header of a dynamic block that is not the last one at bit `bitpos`, with complete codes? (Random bits
    // pass with a probability far below 1e-9; a false positive costs time only: see pgzip.cpp.) `end` bounds the input;
    // 600 bytes behind it must be readable.
    bool probe_dynamic_header(const uint8_t* base, const uint8_t* end, uint64_t bitpos);
    struct MarkerRun {
        bool ok = false;           // the data decoded without an error up to end_bit
        bool ended_final = false;  // end_bit is the end of the final block of a member
        uint64_t end_bit = 0;      // a block boundary: the first one at or behind stop_bit, or the end of the final block
        size_t n_out = 0;          // entries of `out` in use: 32768 place holders of the window + the symbols decoded
    };
    // Decodes from the block header at start_bit, block after block, WITHOUT knowing the 32 KiB in front of it: the
    // output is 16-bit, a value below 256 is a byte, 256 + j stands for byte j of the unknown window (j = 32767: the
    // byte just before the start). `out` only ever grows (a buffer that goes round is not filled with zeros again);
    // n_out of the result says how much of it is in use. Fails (ok = false) on invalid data, at the end of the input
    // inside a block, or beyond max_symbols.
    MarkerRun run_markers(const uint8_t* base, const uint8_t* end, uint64_t start_bit, uint64_t stop_bit, std::vector<uint16_t>* out,
                          size_t max_symbols);
    // Decodes from [*in, in_end) into [*out, out_end). Bytes from out_base on are the history matches may refer to
    // (at least the last 32 KiB produced so far, or everything if less). The 16 bytes behind in_end must be
    // readable (their value does not matter). Nothing is written at or behind out_end.
    Status run(const uint8_t** in, const uint8_t* in_end, bool in_final, const uint8_t* out_base, uint8_t** out, uint8_t* out_end);

private:
    enum State { kBlockHeader, kStored, kHuffman, kDone };
    static constexpr int kLitlenBits = 11, kDistBits = 8;
    static constexpr int kLitlenEntries = (1 << kLitlenBits) + 1400, kDistEntries = (1 << kDistBits) + 700;

    Status run_impl(const uint8_t** in, const uint8_t* in_end, bool in_final, const uint8_t* out_base, uint8_t** out, uint8_t* out_end);
    Status run_generic(const uint8_t** in, const uint8_t* in_end, bool in_final, const uint8_t* out_base, uint8_t** out, uint8_t* out_end);
    Status run_bmi2(const uint8_t** in, const uint8_t* in_end, bool in_final, const uint8_t* out_base, uint8_t** out, uint8_t* out_end);
    MarkerRun run_markers_impl(const uint8_t* base, const uint8_t* end, uint64_t start_bit, uint64_t stop_bit, std::vector<uint16_t>* out,
                               size_t max_symbols);
    MarkerRun run_markers_bmi2(const uint8_t* base, const uint8_t* end, uint64_t start_bit, uint64_t stop_bit, std::vector<uint16_t>* out,
                               size_t max_symbols);
    bool read_dynamic_header(const uint8_t*& in, const uint8_t* in_end, bool in_final);
    void use_fixed_codes();
    static bool build_table(uint32_t* table, int table_bits, int table_cap, const uint8_t* lens, int n_syms, int kind);

    uint64_t bitbuf_;
    unsigned bitcnt_;
    State state_;
    bool last_block_;
    uint32_t stored_left_;
    bool fixed_loaded_;
    bool exact_tail_ = false, stop_at_block_end_ = false;
    uint8_t tail_[2 * 1200 + 64];  // the last bytes of the input, zero padded: a (truncated) block header is parsed from here
    uint32_t litlen_[kLitlenEntries];
    uint32_t dist_[kDistEntries];
};

// CRC-32 as zlib's crc32() computes it (same running value in and out), by carry-less multiplication where the CPU has
// it (PCLMULQDQ: four 128-bit lanes folded over 64 bytes per step), else through zlib.
uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t len);

// The header of a gzip member (RFC 1952) at p: 1 = it is *len bytes long, 0 = more than the n bytes at hand are needed,
// -1 = not a gzip header (magic, method, reserved flags, or its CRC-16 does not match).
int parse_gzip_header(const uint8_t* p, size_t n, size_t* len);

// markers -> bytes: dst[i] = src[i] if it is below 256, else window[src[i] - 256]; `window` has 32768 bytes of which the last
// `window_valid` exist. False if a marker points in front of them (a distance too far back).
bool resolve_markers(const uint16_t* src, size_t n, const uint8_t* window, size_t window_valid, uint8_t* dst);

// One-shot helper (BGZF blocks): the whole stream is in [in, in + in_len), the output has exactly out_len bytes.
bool inflate_exact(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);

}  // namespace mkh
