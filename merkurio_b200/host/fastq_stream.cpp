#include "fastq_stream.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "codecs.h"

namespace mkh {

namespace {

// The input's bytes, whatever its compression (codecs.h).
class RawSource {
public:
    explicit RawSource(const std::string& path) : in_(InputStream::open(path)) {}
    size_t read(char* dst, size_t n) { return in_->read(dst, n); }

private:
    std::unique_ptr<InputStream> in_;
};

struct Line {
    uint32_t off, len;  // without line break and without a trailing '\r'
    bool cr, nl;        // had "\r" before the break / ended with '\n'
    uint32_t next;      // offset of the following line
};

// The line starting at p, where nl[k] is the first line break at or after p (nl holds every '\n' of
// d[0, len)); k moves past the line. Returns false if the line is not complete yet (no '\n' and more
// input may follow), or if there is no line at all (end of data). A last line without '\n' counts only
// when it is not empty (ByteSource::getline has the same rule).
inline bool take_line(const char* d, size_t len, const OffsetList& nl, size_t& k, size_t p, bool eof, Line* out) {
    if (p >= len) return false;
    size_t e;
    if (k < nl.size()) {
        e = nl[k++];
        out->nl = true;
        out->next = (uint32_t)(e + 1);
    } else {
        if (!eof) return false;
        e = len;
        out->nl = false;
        out->next = (uint32_t)len;
    }
    out->cr = e > p && d[e - 1] == '\r';
    out->off = (uint32_t)p;
    out->len = (uint32_t)(e - p - (out->cr ? 1 : 0));
    return true;
}

}  // namespace

bool looks_like_fastq(const std::string& path) {
    try {
        RawSource src(path);
        char c = 0;
        return src.read(&c, 1) == 1 && c == '@';
    } catch (const Error&) {
        return false;
    }
}

struct FastqChunkReader::Shared {
    std::mutex mu;
    std::vector<std::unique_ptr<Chunk>> free_list;
};

FastqChunkReader::FastqChunkReader(const std::string& path, size_t chunk_bytes, size_t depth)
    : path_(path), chunk_bytes_(std::max<size_t>(chunk_bytes, 4096)), depth_(std::max<size_t>(depth, 1)), pool_(new Shared) {
    blocks_.reset(new BlockReader(path_, chunk_bytes_, kHead, 3, true));  // a missing file fails here, in the caller's thread
    thread_ = std::thread([this] { run(); });
}

FastqChunkReader::~FastqChunkReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
    blocks_.reset();
}

std::shared_ptr<Chunk> FastqChunkReader::next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !ready_.empty() || done_; });
    if (!ready_.empty()) {
        std::shared_ptr<Chunk> c = std::move(ready_.front());
        ready_.pop_front();
        lk.unlock();
        cv_.notify_all();
        return c;
    }
    if (!io_error_.empty()) throw Error(io_error_);
    return nullptr;
}

// Second stage: index the 4-line records of each block. Bytes after the last whole record of a block are
// carried over to the front of the next one.
void FastqChunkReader::run() {
    try {
        std::vector<char> carry;
        std::shared_ptr<Shared> pool = pool_;
        BlockReader::Block rb;
        for (bool eof = false; !eof;) {
            const double t_w0 = steady_seconds();
            if (!blocks_->next(rb)) break;  // (an I/O error is thrown by next() after the blocks before it)
            const double t_i0 = steady_seconds();
            t_starved_ += t_i0 - t_w0;
            std::unique_ptr<Chunk> up;  // a recycled chunk, or a new one
            {
                std::lock_guard<std::mutex> lk(pool->mu);
                if (!pool->free_list.empty()) { up = std::move(pool->free_list.back()); pool->free_list.pop_back(); }
            }
            if (!up) up.reset(new Chunk);
            Chunk* c = up.get();
            c->data.swap(rb.data);  // the chunk's previous buffer goes back to the reader with the next call
            eof = rb.last;
            c->recs.clear();
            c->nl.clear();
            c->failed = false;
            size_t begin, have;
            // the reading stage has located the line breaks of its block already (BlockReader scan_lines); only those of
            // the bytes carried over are missing — unless the block had to be moved (an over-long record)
            const bool pre = rb.has_nl && carry.size() <= kHead;
            if (carry.size() <= kHead) {
                begin = kHead - carry.size();
                have = kHead + rb.n;
                if (!carry.empty()) std::memcpy(c->data.data() + begin, carry.data(), carry.size());
            } else {
                // a record longer than the head room (offsets are 32-bit)
                if (carry.size() > ((size_t)1 << 30)) throw Error("FASTQ record larger than 1 GiB");
                ByteBuf joined(kHead + carry.size() + std::max(rb.n, chunk_bytes_));
                std::memcpy(joined.data() + kHead, carry.data(), carry.size());
                std::memcpy(joined.data() + kHead + carry.size(), c->data.data() + kHead, rb.n);
                c->data.swap(joined);
                begin = kHead;
                have = kHead + carry.size() + rb.n;
            }
            carry.clear();
            // index the whole records of data[begin, have): line breaks are located a stretch at a time and the
            // records of the stretch parsed while it is still in the cache
            const char* d = c->data.data();
            const OffsetList& nl = c->nl;
            size_t p = begin, k = 0;  // nl[k]: the first line break at or after p
            bool stuck = false;       // malformed or truncated: nothing more to parse in this block
            size_t n_head_breaks = 0;  // line breaks of the bytes carried over: c->nl[n_head_breaks + i] == rb.nl[i]
            if (pre) {
                find_line_breaks(d, begin, kHead, c->nl);
                n_head_breaks = c->nl.size();
                c->nl.append(rb.nl);
            }
            for (size_t scanned = begin; !stuck && (scanned < have || (eof && p < have));) {
                const size_t upto = pre ? have : std::min(have, scanned + kStretch);
                if (!pre) find_line_breaks(d, scanned, upto, c->nl);
                scanned = upto;
                const bool last = eof && scanned == have;  // only then a line without '\n' is complete
                for (;;) {
                    // The common record — "@id\nseq\n+\nqual\n", no '\r' — is recognised from the offsets of its line
                    // breaks and from what the reading stage noted beside them (BlockReader::kNl*), without touching
                    // the block's bytes: they have left this core's caches, and pulling them through once more was
                    // the whole cost of indexing. Everything else takes the general rules below, which give such a
                    // record the same span.
                    if (pre) {
                        using BR = BlockReader;
                        const uint8_t* cx = rb.nl_ctx.data();
                        while (k > n_head_breaks && k + 4 <= nl.size()) {
                            const size_t j = k - n_head_breaks;  // cx[j] belongs to nl[k]; nl[k - 1] + 1 == p
                            const size_t e0 = nl[k], e1 = nl[k + 1], e2 = nl[k + 2], e3 = nl[k + 3];
                            if ((cx[j - 1] & (BR::kNlAt | BR::kNlNoNext)) != BR::kNlAt || (cx[j] & (BR::kNlCr | BR::kNlNoPrev)) ||
                                (cx[j + 1] & (BR::kNlCr | BR::kNlNoPrev | BR::kNlPlus | BR::kNlNoNext)) != BR::kNlPlus || e2 != e1 + 2 ||
                                (cx[j + 3] & (BR::kNlCr | BR::kNlNoPrev)) || e1 - e0 != e3 - e2)
                                break;
                            RecSpan r;
                            r.start = (uint32_t)p;
                            r.id_len = (uint32_t)(e0 - p - 1);
                            r.seq_off = (uint32_t)(e0 + 1);
                            r.seq_len = (uint32_t)(e1 - e0 - 1);
                            r.qual_off = (uint32_t)(e2 + 1);
                            r.end = (uint32_t)(e3 + 1);
                            r.crlf = 0;
                            r.plain = 1;
                            c->recs.push_back(r);
                            p = e3 + 1;
                            k += 4;
                        }
                    }
                    Line h;
                    const size_t kp = k;
                    size_t q = p, kh = k;
                    bool got = false;
                    for (;;) {  // blank lines before a record
                        kh = k;
                        got = take_line(d, scanned, nl, k, q, last, &h);
                        if (!got || h.len != 0) break;
                        q = h.next;
                    }
                    if (!got) {
                        if (last) p = have;  // only blank lines were left
                        else k = kp;
                        break;
                    }
                    Line s, pl, ql;
                    bool complete = take_line(d, scanned, nl, k, h.next, last, &s) && take_line(d, scanned, nl, k, s.next, last, &pl) &&
                                    take_line(d, scanned, nl, k, pl.next, last, &ql);
                    if (!complete) {
                        if (last) c->failed = stuck = true;  // truncated record
                        else {                               // resumed (or carried over) from this header on
                            p = q;
                            k = kh;
                        }
                        break;
                    }
                    if (d[h.off] != '@' || pl.len == 0 || d[pl.off] != '+' || s.len != ql.len) { c->failed = stuck = true; break; }
                    RecSpan r;
                    r.start = h.off;
                    r.id_len = h.len - 1;
                    r.seq_off = s.off;
                    r.seq_len = s.len;
                    r.qual_off = ql.off;
                    r.end = ql.next;
                    r.crlf = h.cr;
                    r.plain = !h.cr && !s.cr && !pl.cr && !ql.cr && pl.len == 1 && ql.nl;
                    c->recs.push_back(r);
                    p = ql.next;
                }
                if (last) break;
            }
            if (c->failed) eof = true;  // nothing after a malformed record is looked at
            else if (p < have) carry.assign(d + p, d + have);
            c->len = p;
            t_index_ += steady_seconds() - t_i0;
            if (c->recs.empty() && !c->failed) {
                std::lock_guard<std::mutex> lk(pool->mu);
                pool->free_list.push_back(std::move(up));
                continue;
            }
            // hand the chunk over; when the last reference drops, its buffers go back to the free list
            std::shared_ptr<Chunk> sp(up.release(), [pool](Chunk* ch) {
                std::unique_ptr<Chunk> back(ch);
                std::lock_guard<std::mutex> lk(pool->mu);
                if (pool->free_list.size() < 16) pool->free_list.push_back(std::move(back));
            });
            const double t_b0 = steady_seconds();
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return ready_.size() < depth_ || stop_; });
            t_blocked_ += steady_seconds() - t_b0;
            if (stop_) return;
            ready_.push_back(std::move(sp));
            lk.unlock();
            cv_.notify_all();
        }
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu_);
        if (io_error_.empty()) io_error_ = e.what();
    }
    if (std::getenv("MERKURIO_TIMING"))
        std::fprintf(stderr, "[merkurio] FASTQ reader %s: read %.3f s | index %.3f s, waiting for input %.3f s, blocked on the packer %.3f s\n",
                     path_.c_str(), blocks_->seconds_reading(), t_index_, t_starved_, t_blocked_);
    {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
    }
    cv_.notify_all();
}

}  // namespace mkh
