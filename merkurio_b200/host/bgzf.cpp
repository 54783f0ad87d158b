#include "bgzf.h"

#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "inflate.h"

namespace mkh {

namespace {
constexpr size_t kInChunk = 8u << 20;  // compressed bytes inflated per round

// Size of the BGZF block starting at p (n bytes available), 0 if the header is incomplete, -1 if it
// is not a BGZF header.
long bgzf_block_size(const unsigned char* p, size_t n) {
    if (n < 18) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return -1;
    const size_t xlen = p[10] | ((size_t)p[11] << 8);
    if (n < 12 + xlen) return 0;
    for (size_t q = 12; q + 4 <= 12 + xlen;) {
        const size_t slen = p[q + 2] | ((size_t)p[q + 3] << 8);
        if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) return (long)(p[q + 4] | ((size_t)p[q + 5] << 8)) + 1;
        q += 4 + slen;
    }
    return -1;
}
}  // namespace

bool is_bgzf(int fd) {
    unsigned char h[64];
    ssize_t n = ::pread(fd, h, sizeof h, 0);
    return n >= 18 && bgzf_block_size(h, (size_t)n) > 0;
}

BgzfReader::BgzfReader(int fd, int n_threads) : fd_(fd) {
    ::lseek(fd_, 0, SEEK_SET);
    in_.resize(kInChunk + (1u << 16) + 64);
    for (int t = 1; t < std::max(n_threads, 1); ++t) threads_.emplace_back([this] { worker(); });
}

BgzfReader::~BgzfReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : threads_) t.join();
    if (fd_ >= 0) ::close(fd_);
}

void BgzfReader::inflate_block(const Block& b) {
    const unsigned char* p = in_.data() + b.in_off;
    const size_t xlen = p[10] | ((size_t)p[11] << 8);
    const size_t hdr = 12 + xlen;
    if (b.in_len < hdr + 8) throw Error("corrupt BGZF block");
    if (b.out_len == 0) return;
    static const bool use_zlib = std::getenv("MERKURIO_ZLIB_INFLATE") != nullptr;  // the round-1 path, kept for comparison
    if (use_zlib) {
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) throw Error("inflateInit2 failed");
        zs.next_in = const_cast<unsigned char*>(p + hdr);
        zs.avail_in = (unsigned)(b.in_len - hdr - 8);
        zs.next_out = reinterpret_cast<unsigned char*>(out_.data() + b.out_off);
        zs.avail_out = (unsigned)b.out_len;
        const int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END || zs.avail_out != 0) throw Error("Error while decompressing the input");
    } else if (!inflate_exact(p + hdr, b.in_len - hdr - 8, reinterpret_cast<uint8_t*>(out_.data() + b.out_off), b.out_len)) {
        // (the 8 trailer bytes and the next block, or the padding of in_, follow the payload: readable)
        throw Error("Error while decompressing the input");
    }
    const unsigned char* t = p + b.in_len - 8;
    const uint32_t want = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    const uint32_t got = crc32_fast(0, reinterpret_cast<const uint8_t*>(out_.data() + b.out_off), b.out_len);
    if (want != got) throw Error("Error while decompressing the input (CRC mismatch)");
}

void BgzfReader::worker() {
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
        }
        std::string err;
        size_t bad = SIZE_MAX;
        for (size_t i; (i = next_.fetch_add(1)) < blocks_.size();) {
            try {
                inflate_block(blocks_[i]);
            } catch (const std::exception& e) {
                if (i < bad) { bad = i; err = e.what(); }
            }
        }
        std::lock_guard<std::mutex> lk(mu_);
        if (bad < first_bad_) { first_bad_ = bad; error_ = err; }
        ++finished_workers_;
        cv_done_.notify_all();
    }
}

// Read the next stretch of whole blocks and inflate them (all threads, this one included).
bool BgzfReader::refill() {
    out_pos_ = out_len_ = 0;
    // the blocks in front of a broken one were handed out by the round before
    if (!pending_error_.empty()) throw Error(pending_error_);
    while (!eof_ && in_have_ < kInChunk) {
        ssize_t n = ::read(fd_, in_.data() + in_have_, in_.size() - in_have_);
        if (n < 0) {
            if (errno == EINTR) continue;
            throw Error(std::string("read failed: ") + std::strerror(errno));
        }
        if (n == 0) eof_ = true;
        in_have_ += (size_t)n;
    }
    blocks_.clear();
    size_t p = 0, out_total = 0;
    while (p < in_have_) {
        long bs = bgzf_block_size(in_.data() + p, in_have_ - p);
        // a broken block: the whole blocks in front of it are still inflated and handed out (the reference reads
        // block by block and fails at the record it cannot read), the error follows with the next round
        const char* broken = nullptr;
        if (bs < 0) broken = "Error while decompressing the input (not a BGZF block)";
        else if (bs == 0 || p + (size_t)bs > in_have_) {
            if (!eof_) break;
            broken = "unexpected end of file";
        }
        size_t isize = 0;
        if (!broken) {
            const unsigned char* t = in_.data() + p + bs - 4;
            isize = t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
            if (isize > (1u << 16)) broken = "corrupt BGZF block";
        }
        if (broken) {
            if (blocks_.empty()) throw Error(broken);
            pending_error_ = broken;
            break;
        }
        blocks_.push_back(Block{p, (size_t)bs, out_total, isize});
        out_total += isize;
        p += (size_t)bs;
    }
    if (blocks_.empty()) return false;
    if (out_.size() < out_total) out_.resize(out_total);
    next_.store(0);
    {
        std::lock_guard<std::mutex> lk(mu_);
        finished_workers_ = 0;
        first_bad_ = SIZE_MAX;
        error_.clear();
        ++generation_;
    }
    cv_work_.notify_all();
    std::string err;
    size_t bad = SIZE_MAX;
    for (size_t i; (i = next_.fetch_add(1)) < blocks_.size();) {
        try {
            inflate_block(blocks_[i]);
        } catch (const std::exception& e) {
            if (i < bad) { bad = i; err = e.what(); }
        }
    }
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return finished_workers_ == threads_.size(); });
        if (first_bad_ < bad) { bad = first_bad_; err = error_; }
    }
    if (bad != SIZE_MAX) {
        // the first block that does not inflate: what precedes it is good
        if (blocks_[bad].out_off == 0) throw Error(err);
        pending_error_ = err;
        out_total = blocks_[bad].out_off;
    }
    // keep the partial block at the end for the next round
    std::memmove(in_.data(), in_.data() + p, in_have_ - p);
    in_have_ -= p;
    out_len_ = out_total;
    return true;
}

size_t BgzfReader::read(char* dst, size_t n) {
    size_t got = 0;
    while (got < n) {
        if (out_pos_ == out_len_) {
            // a round may hold only empty blocks (the EOF marker): keep going until data or the end
            bool more = true;
            while (out_pos_ == out_len_ && (more = refill())) {}
            if (!more) break;
        }
        size_t take = std::min(n - got, out_len_ - out_pos_);
        std::memcpy(dst + got, out_.data() + out_pos_, take);
        out_pos_ += take;
        got += take;
        if (got) break;  // hand back what is there; the caller asks again
    }
    return got;
}

}  // namespace mkh
