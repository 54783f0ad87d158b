#include "fastq_stream.h"

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

// The line starting at p. Returns false if it is not complete yet (no '\n' and more input may
// follow), or if there is no line at all (end of data). A last line without '\n' counts only when it
// is not empty (ByteSource::getline has the same rule).
inline bool take_line(const char* d, size_t len, size_t p, bool eof, Line* out) {
    if (p >= len) return false;
    const char* nl = static_cast<const char*>(std::memchr(d + p, '\n', len - p));
    size_t e;
    if (nl) {
        e = (size_t)(nl - d);
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
    // open here so that a missing file fails in the caller's thread, with the caller's context
    { RawSource probe(path_); }
    thread_ = std::thread([this] { run(); });
}

FastqChunkReader::~FastqChunkReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
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

void FastqChunkReader::run() {
    try {
        RawSource src(path_);
        std::vector<char> carry;
        bool eof = false;
        std::shared_ptr<Shared> pool = pool_;
        while (!eof) {
            // a recycled buffer, or a new one
            std::unique_ptr<Chunk> up;
            {
                std::lock_guard<std::mutex> lk(pool->mu);
                if (!pool->free_list.empty()) { up = std::move(pool->free_list.back()); pool->free_list.pop_back(); }
            }
            if (!up) up.reset(new Chunk);
            Chunk* c = up.get();
            c->recs.clear();
            c->failed = false;
            if (c->data.size() < chunk_bytes_ + carry.size()) c->data.resize(chunk_bytes_ + carry.size());
            size_t have = carry.size();
            if (have) std::memcpy(c->data.data(), carry.data(), have);
            carry.clear();
            size_t consumed = 0;
            for (;;) {
                while (!eof && have < c->data.size()) {
                    size_t n = src.read(c->data.data() + have, c->data.size() - have);
                    if (n == 0) eof = true;
                    have += n;
                }
                // index the whole records of data[0, have)
                const char* d = c->data.data();
                size_t p = consumed;
                for (;;) {
                    Line h;
                    size_t q = p;
                    bool got = false;
                    while ((got = take_line(d, have, q, eof, &h)) && h.len == 0) q = h.next;  // blank lines before a record
                    if (!got) {
                        if (eof) p = have;  // only blank lines were left
                        break;
                    }
                    Line s, pl, ql;
                    bool complete = take_line(d, have, h.next, eof, &s) && take_line(d, have, s.next, eof, &pl) &&
                                    take_line(d, have, pl.next, eof, &ql);
                    if (!complete) {
                        if (eof) c->failed = true;  // truncated record
                        else p = q;                 // resume at this header once more input is here
                        break;
                    }
                    if (d[h.off] != '@' || pl.len == 0 || d[pl.off] != '+' || s.len != ql.len) { c->failed = true; break; }
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
                consumed = p;
                if (c->failed || eof || !c->recs.empty()) break;
                // not even one whole record in a full buffer: grow it and keep reading (offsets are 32-bit)
                if (c->data.size() > ((size_t)1 << 30)) throw Error("FASTQ record larger than 1 GiB");
                c->data.resize(c->data.size() * 2);
            }
            if (c->failed) eof = true;  // nothing after a malformed record is looked at
            else if (consumed < have) carry.assign(c->data.data() + consumed, c->data.data() + have);
            c->len = consumed;
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
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return ready_.size() < depth_ || stop_; });
            if (stop_) return;
            ready_.push_back(std::move(sp));
            lk.unlock();
            cv_.notify_all();
        }
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu_);
        io_error_ = e.what();
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
    }
    cv_.notify_all();
}

}  // namespace mkh
