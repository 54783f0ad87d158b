#include "aln_stream.h"

#include <cstring>

namespace mkh {

struct AlnChunkReader::Shared {
    std::mutex mu;
    std::vector<std::unique_ptr<AlnChunk>> free_list;
};

AlnChunkReader::AlnChunkReader(std::unique_ptr<AlnReader> reader, size_t chunk_bytes, size_t depth)
    : reader_(std::move(reader)), chunk_bytes_(std::max<size_t>(chunk_bytes, 4096)), depth_(std::max<size_t>(depth, 1)), pool_(new Shared) {
    thread_ = std::thread([this] { run(); });
}

AlnChunkReader::~AlnChunkReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
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

// Index the SAM lines of d[p, have). Returns the offset of the first byte not consumed.
size_t index_sam(AlnChunk* c, size_t p, size_t have, bool eof) {
    const char* d = c->data.data();
    while (p < have) {
        const char* nl = static_cast<const char*>(std::memchr(d + p, '\n', have - p));
        size_t e;
        if (nl) e = (size_t)(nl - d);
        else if (eof) e = have;
        else break;
        size_t next = nl ? e + 1 : have;
        size_t le = e;
        if (le > p && d[le - 1] == '\r') --le;
        if (le == p) { p = next; continue; }  // blank line
        // fields: QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL [tags]
        size_t starts[12];
        int nf = 0;
        starts[nf++] = p;
        for (size_t q = p; nf < 12;) {
            const char* tab = static_cast<const char*>(std::memchr(d + q, '\t', le - q));
            if (!tab) break;
            q = (size_t)(tab - d) + 1;
            starts[nf++] = q;
        }
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

void AlnChunkReader::run() {
    try {
        ByteSource& src = reader_->source();
        const bool bam = reader_->is_bam();
        std::vector<char> carry;
        bool eof = false;
        std::shared_ptr<Shared> pool = pool_;
        while (!eof) {
            std::unique_ptr<AlnChunk> up;
            {
                std::lock_guard<std::mutex> lk(pool->mu);
                if (!pool->free_list.empty()) { up = std::move(pool->free_list.back()); pool->free_list.pop_back(); }
            }
            if (!up) up.reset(new AlnChunk);
            AlnChunk* c = up.get();
            c->recs.clear();
            c->error.clear();
            c->bam = bam;
            if (c->data.size() < chunk_bytes_ + carry.size()) c->data.resize(chunk_bytes_ + carry.size());
            size_t have = carry.size();
            if (have) std::memcpy(c->data.data(), carry.data(), have);
            carry.clear();
            size_t consumed = 0;
            for (;;) {
                while (!eof && have < c->data.size()) {
                    size_t n = src.read_some(c->data.data() + have, c->data.size() - have);
                    if (n == 0) eof = true;
                    have += n;
                }
                consumed = bam ? index_bam(c, consumed, have, eof) : index_sam(c, consumed, have, eof);
                if (!c->error.empty() || eof || !c->recs.empty()) break;
                if (c->data.size() > (1u << 30)) throw Error("record larger than 1 GiB");
                c->data.resize(c->data.size() * 2);  // not even one whole record in a full buffer
            }
            if (!c->error.empty()) eof = true;
            else if (consumed < have) carry.assign(c->data.data() + consumed, c->data.data() + have);
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
