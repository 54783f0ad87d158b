#include "aln_stream.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mkh {

struct AlnChunkReader::Shared {
    std::mutex mu;
    std::vector<std::unique_ptr<AlnChunk>> free_list;
};

AlnChunkReader::AlnChunkReader(std::unique_ptr<AlnReader> reader, size_t chunk_bytes, size_t depth)
    : reader_(std::move(reader)), chunk_bytes_(std::max<size_t>(chunk_bytes, 4096)), depth_(std::max<size_t>(depth, 1)), pool_(new Shared) {
    ByteSource* src = &reader_->source();  // what follows the header
    blocks_.reset(new BlockReader([src](char* dst, size_t n) { return src->read_some(dst, n); }, chunk_bytes_, kHead));
    thread_ = std::thread([this] { run(); });
}

AlnChunkReader::~AlnChunkReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
    blocks_.reset();
}

std::shared_ptr<AlnChunk> AlnChunkReader::next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !ready_.empty() || done_; });
    if (!ready_.empty()) {
        std::shared_ptr<AlnChunk> c = std::move(ready_.front());
        ready_.pop_front();
        lk.unlock();
        cv_.notify_all();
        return c;
    }
    if (!io_error_.empty()) throw Error(io_error_);
    return nullptr;
}

namespace {

// Index the SAM lines of d[p, have). Returns the offset of the first byte not consumed. Line breaks and
// tabs are located a stretch at a time (sep: scratch), the lines of the stretch split while it is in the cache.
size_t index_sam(AlnChunk* c, size_t p, size_t have, bool eof, OffsetList& sep) {
    const char* d = c->data.data();
    const size_t kStretch = (size_t)128 << 10;
    sep.clear();
    size_t scanned = p, k = 0;  // sep[k]: the first separator at or after p
    while (p < have) {
        // fields: QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL [tags]
        size_t starts[12];
        int nf = 0;
        starts[nf++] = p;
        size_t kk = k, e = have;
        bool nl = false;
        for (;;) {
            if (kk == sep.n) {
                if (scanned == have) break;
                const size_t upto = std::min(have, scanned + kStretch);
                find_breaks_and_tabs(d, scanned, upto, sep);
                scanned = upto;
                continue;
            }
            const size_t at = sep.p[kk++];
            if (d[at] == '\n') { e = at; nl = true; break; }
            if (nf < 12) starts[nf++] = at + 1;
        }
        if (!nl && !eof) break;  // the line continues in the next block
        k = kk;
        const size_t next = nl ? e + 1 : have;
        size_t le = e;
        if (le > p && d[le - 1] == '\r') --le;
        if (le == p) { p = next; continue; }  // blank line
        if (nf < 11) { c->error = "truncated record"; return p; }
        AlnSpan s;
        s.off = (uint32_t)p;
        s.len = (uint32_t)(le - p);
        s.name_off = (uint32_t)p;
        s.name_len = (uint32_t)(starts[1] - 1 - p);
        s.seq_off = (uint32_t)starts[9];
        s.l_seq = (uint32_t)(starts[10] - 1 - starts[9]);
        if (s.l_seq == 1 && d[s.seq_off] == '*') s.l_seq = 0;
        c->recs.push_back(s);
        p = next;
    }
    return p;
}

// Index the BAM records of d[p, have).
size_t index_bam(AlnChunk* c, size_t p, size_t have, bool eof) {
    const char* d = c->data.data();
    while (p < have) {
        if (have - p < 4) {
            if (eof) c->error = "unexpected end of file";
            break;
        }
        int32_t block_size;
        std::memcpy(&block_size, d + p, 4);
        if (block_size < 32) { c->error = "truncated record"; break; }
        if (have - p - 4 < (size_t)block_size) {
            if (eof) c->error = "unexpected end of file";
            break;
        }
        const char* b = d + p + 4;
        uint8_t l_read_name = (uint8_t)b[8];
        uint16_t n_cigar;
        int32_t l_seq;
        std::memcpy(&n_cigar, b + 12, 2);
        std::memcpy(&l_seq, b + 16, 4);
        size_t seq_at = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar;
        if (l_seq < 0 || seq_at + ((size_t)l_seq + 1) / 2 + (size_t)l_seq > (size_t)block_size) { c->error = "truncated record"; break; }
        AlnSpan s;
        s.off = (uint32_t)(p + 4);
        s.len = (uint32_t)block_size;
        s.name_off = s.off + 32;
        s.name_len = l_read_name ? l_read_name - 1u : 0u;
        s.seq_off = s.off + (uint32_t)seq_at;
        s.l_seq = (uint32_t)l_seq;
        c->recs.push_back(s);
        p += 4 + (size_t)block_size;
    }
    return p;
}

}  // namespace

// The indexing thread: records of each block the reading thread (blocks_) hands over. Bytes after the last
// whole record of a block are carried over to the front of the next one.
void AlnChunkReader::run() {
    double t_index = 0, t_starved = 0, t_blocked = 0;
    try {
        const bool bam = reader_->is_bam();
        std::vector<char> carry;
        std::shared_ptr<Shared> pool = pool_;
        BlockReader::Block rb;
        OffsetList sep;
        for (bool eof = false; !eof;) {
            const double t_w0 = steady_seconds();
            if (!blocks_->next(rb)) break;  // (an I/O error is thrown by next() after the blocks before it)
            const double t_i0 = steady_seconds();
            t_starved += t_i0 - t_w0;
            std::unique_ptr<AlnChunk> up;
            {
                std::lock_guard<std::mutex> lk(pool->mu);
                if (!pool->free_list.empty()) { up = std::move(pool->free_list.back()); pool->free_list.pop_back(); }
            }
            if (!up) up.reset(new AlnChunk);
            AlnChunk* c = up.get();
            c->data.swap(rb.data);  // the chunk's previous buffer goes back to the reader with the next call
            eof = rb.last;
            c->recs.clear();
            c->error.clear();
            c->bam = bam;
            size_t begin, have;
            if (carry.size() <= kHead) {
                begin = kHead - carry.size();
                have = kHead + rb.n;
                if (!carry.empty()) std::memcpy(c->data.data() + begin, carry.data(), carry.size());
            } else {
                // a record longer than the head room (offsets are 32-bit)
                if (carry.size() > ((size_t)1 << 30)) throw Error("record larger than 1 GiB");
                ByteBuf joined(kHead + carry.size() + std::max(rb.n, chunk_bytes_));
                std::memcpy(joined.data() + kHead, carry.data(), carry.size());
                std::memcpy(joined.data() + kHead + carry.size(), c->data.data() + kHead, rb.n);
                c->data.swap(joined);
                begin = kHead;
                have = kHead + carry.size() + rb.n;
            }
            carry.clear();
            const size_t consumed = bam ? index_bam(c, begin, have, eof) : index_sam(c, begin, have, eof, sep);
            if (!c->error.empty()) eof = true;
            else if (consumed < have) carry.assign(c->data.data() + consumed, c->data.data() + have);
            t_index += steady_seconds() - t_i0;
            if (c->recs.empty() && c->error.empty()) {
                std::lock_guard<std::mutex> lk(pool->mu);
                pool->free_list.push_back(std::move(up));
                continue;
            }
            std::shared_ptr<AlnChunk> sp(up.release(), [pool](AlnChunk* ch) {
                std::unique_ptr<AlnChunk> back(ch);
                std::lock_guard<std::mutex> lk(pool->mu);
                if (pool->free_list.size() < 24) pool->free_list.push_back(std::move(back));
            });
            const double t_b0 = steady_seconds();
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return ready_.size() < depth_ || stop_; });
            t_blocked += steady_seconds() - t_b0;
            if (stop_) return;
            ready_.push_back(std::move(sp));
            lk.unlock();
            cv_.notify_all();
        }
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu_);
        io_error_ = e.what();
    }
    if (std::getenv("MERKURIO_TIMING"))
        std::fprintf(stderr, "[merkurio] SAM/BAM reader: read %.3f s | index %.3f s, waiting for input %.3f s, blocked on the packer %.3f s\n",
                     blocks_->seconds_reading(), t_index, t_starved, t_blocked);
    {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
    }
    cv_.notify_all();
}

}  // namespace mkh
