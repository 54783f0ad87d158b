#include "fasta_pipeline.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "codecs.h"

namespace mkh {

// ------------------------------------------------------------------------------------------------
struct FastaChunkReader::Shared {
    std::mutex mu;
    std::vector<std::unique_ptr<FaChunk>> free_list;
};

FastaChunkReader::FastaChunkReader(const std::string& path, size_t chunk_bytes, size_t depth)
    : path_(path), chunk_bytes_(std::max<size_t>(chunk_bytes, 4096)), depth_(std::max<size_t>(depth, 1)), pool_(new Shared) {
    blocks_.reset(new BlockReader(path_, chunk_bytes_, kHead));  // a missing file fails here, in the caller's thread
    thread_ = std::thread([this] { run(); });
}

FastaChunkReader::~FastaChunkReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
    blocks_.reset();
}

std::shared_ptr<FaChunk> FastaChunkReader::next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !ready_.empty() || done_; });
    if (!ready_.empty()) {
        std::shared_ptr<FaChunk> c = std::move(ready_.front());
        ready_.pop_front();
        lk.unlock();
        cv_.notify_all();
        return c;
    }
    if (!io_error_.empty()) throw Error(io_error_);
    return nullptr;
}

// The indexing thread: lines of each block the reading thread (blocks_) hands over. An unfinished last
// line is carried over to the front of the next block.
void FastaChunkReader::run() {
    double t_index = 0, t_starved = 0, t_blocked = 0;
    try {
        std::vector<char> carry;
        std::shared_ptr<Shared> pool = pool_;
        BlockReader::Block rb;
        OffsetList nl;
        for (bool eof = false; !eof;) {
            const double t_w0 = steady_seconds();
            if (!blocks_->next(rb)) break;  // (an I/O error is thrown by next() after the blocks before it)
            const double t_i0 = steady_seconds();
            t_starved += t_i0 - t_w0;
            std::unique_ptr<FaChunk> up;
            {
                std::lock_guard<std::mutex> lk(pool->mu);
                if (!pool->free_list.empty()) { up = std::move(pool->free_list.back()); pool->free_list.pop_back(); }
            }
            if (!up) up.reset(new FaChunk);
            FaChunk* c = up.get();
            c->data.swap(rb.data);  // the chunk's previous buffer goes back to the reader with the next call
            eof = rb.last;
            c->lines.clear();
            size_t begin, have;
            if (carry.size() <= kHead) {
                begin = kHead - carry.size();
                have = kHead + rb.n;
                if (!carry.empty()) std::memcpy(c->data.data() + begin, carry.data(), carry.size());
            } else {
                // a line longer than the head room (offsets are 32-bit)
                if (carry.size() > ((size_t)1 << 30)) throw Error("FASTA line longer than 1 GiB");
                ByteBuf joined(kHead + carry.size() + std::max(rb.n, chunk_bytes_));
                std::memcpy(joined.data() + kHead, carry.data(), carry.size());
                std::memcpy(joined.data() + kHead + carry.size(), c->data.data() + kHead, rb.n);
                c->data.swap(joined);
                begin = kHead;
                have = kHead + carry.size() + rb.n;
            }
            carry.clear();
            // line breaks are located a stretch at a time, the lines noted while the stretch is still in the cache
            const char* d = c->data.data();
            size_t p = begin;
            for (size_t scanned = begin; scanned < have;) {
                const size_t upto = std::min(have, scanned + kStretch);
                nl.clear();
                find_line_breaks(d, scanned, upto, nl);
                scanned = upto;
                const size_t base = c->lines.size();
                c->lines.resize(base + nl.n);
                FaLine* w = c->lines.data() + base;
                for (size_t i = 0; i < nl.n; ++i) {
                    const size_t e = nl.p[i];
                    w[i] = FaLine{(uint32_t)p, (uint32_t)(e - p), (uint8_t)(e > p && d[p] == '>')};
                    p = e + 1;
                }
            }
            if (eof && p < have) {  // an unterminated last line (it is not empty)
                c->lines.push_back(FaLine{(uint32_t)p, (uint32_t)(have - p), (uint8_t)(d[p] == '>')});
                p = have;
            }
            if (p < have) carry.assign(d + p, d + have);
            c->last = eof;
            t_index += steady_seconds() - t_i0;
            if (c->lines.empty()) {
                std::lock_guard<std::mutex> lk(pool->mu);
                pool->free_list.push_back(std::move(up));
                continue;
            }
            std::shared_ptr<FaChunk> sp(up.release(), [pool](FaChunk* ch) {
                std::unique_ptr<FaChunk> back(ch);
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
        std::fprintf(stderr, "[merkurio] FASTA reader %s: read %.3f s | index %.3f s, waiting for input %.3f s, blocked on the packer %.3f s\n",
                     path_.c_str(), blocks_->seconds_reading(), t_index, t_starved, t_blocked);
    {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = true;
    }
    cv_.notify_all();
}

bool looks_like_fasta(const std::string& path) {
    try {
        std::unique_ptr<InputStream> src = InputStream::open(path);
        char buf[4096];
        for (;;) {
            size_t n = src->read(buf, sizeof buf);
            if (n == 0) return false;
            for (size_t i = 0; i < n; ++i) {
                if (buf[i] == '\n' || buf[i] == '\r') continue;  // blank lines before the first record
                return buf[i] == '>';
            }
        }
    } catch (const Error&) {
        return false;
    }
}

// ------------------------------------------------------------------------------------------------
FastaPipeline::FastaPipeline(EngineSet& engines, std::unique_ptr<FastaChunkReader> reader, mk_mode mode, bool keep_text, BatchConsumer consumer)
    : SlotPipeline(engines, MK_ENC_ASCII, mode, std::move(consumer)), rd_(std::move(reader)), keep_text_(keep_text) {
    overlap_ = es_.max_pattern_len ? es_.max_pattern_len - 1 : 0;
}

FastaPipeline::~FastaPipeline() { stop_packer(); }

void FastaPipeline::begin() { cur_ = rd_->next(); }

void FastaPipeline::flush_range() {
    if (range_open_ && rec_ && keep_text_) rec_->raw.push_back(FaRecord::Range{cur_, range_off_, range_end_ - range_off_});
    range_open_ = false;
}

bool FastaPipeline::fill(PackedBatch& b) {
    b.n_records = 0;
    b.n_units = b.n_bytes = b.total_bases = 0;
    b.seg[0].clear();
    b.seg[1].clear();
    b.error_chain.clear();
    std::shared_ptr<FaBatchInfo> info(new FaBatchInfo);
    b.extra = info;
    if (input_done_) return false;
    const uint64_t cap = es_.max_bytes;
    const uint32_t max_rec = es_.max_records;
    // open a piece of rec_ in this batch; false if the batch has no record slot left
    auto open_piece = [&](bool first) {
        if (b.n_records >= max_rec) return false;
        uint64_t lead = first ? 0 : std::min<uint64_t>(overlap_, rec_->len);
        if (lead) {  // the record's last bases again: they are the tail of the slot filled before this one
            std::memmove(b.seq + b.n_bytes, prev_seq_ + prev_bytes_ - lead, lead);
        }
        b.off[b.n_records] = b.n_bytes;
        info->pieces.push_back(FaPiece{rec_, first, false, rec_->len - lead, (uint32_t)lead});
        b.n_bytes += lead;
        b.n_records += 1;
        rec_open_piece_ = true;
        return true;
    };
    auto end_record = [&] {  // rec_'s piece is the last one of this batch
        flush_range();
        info->pieces.back().last = true;
        rec_.reset();
        rec_open_piece_ = false;
    };
    for (;;) {
        while (cur_ && line_ == cur_->lines.size()) {
            flush_range();
            cur_ = rd_->next();
            line_ = 0;
            line_pos_ = 0;
        }
        if (!cur_) {  // end of input
            if (rec_) {
                if (!rec_open_piece_ && !open_piece(false)) break;
                end_record();
            }
            input_done_ = true;
            break;
        }
        const FaLine& ln = cur_->lines[line_];
        const char* text = cur_->data.data() + ln.off;
        const bool cr = ln.len && text[ln.len - 1] == '\r';
        if (!started_) {
            if (ln.len == (cr ? 1u : 0u)) { ++line_; continue; }  // blank lines before the first record
            started_ = true;
        }
        if (ln.header) {
            if (rec_) {  // the previous record ends here
                if (!rec_open_piece_ && !open_piece(false)) break;
                end_record();
            }
            if (b.n_records >= max_rec) break;  // the header opens the next batch
            rec_.reset(new FaRecord);
            rec_->id.assign(text + 1, ln.len - 1 - (cr ? 1 : 0));
            rec_->crlf = cr;
            open_piece(true);
            ++line_;
            continue;
        }
        // a sequence line of rec_
        if (!rec_open_piece_) {
            if (b.n_bytes + std::min<uint64_t>(overlap_, rec_->len) >= cap && b.n_records > 0) break;
            if (!open_piece(false)) break;
        }
        if (line_pos_ == 0) {
            if (!range_open_) { range_open_ = true; range_off_ = ln.off; }
            range_end_ = ln.off + ln.len;
        }
        const uint32_t bases = ln.len - (cr ? 1 : 0);
        const uint64_t room = cap - b.n_bytes;
        const uint64_t take = std::min<uint64_t>(bases - line_pos_, room);
        if (take) {
            std::memcpy(b.seq + b.n_bytes, text + line_pos_, take);
            b.n_bytes += take;
            rec_->len += take;
            line_pos_ += (uint32_t)take;
        }
        if (line_pos_ == bases) {
            ++line_;
            line_pos_ = 0;
            continue;
        }
        // the slot is full in the middle of the record: it goes on in the next batch
        rec_open_piece_ = false;
        break;
    }
    b.off[b.n_records] = b.n_bytes;
    b.n_units = b.total_bases = b.n_bytes;
    prev_seq_ = b.seq;
    prev_bytes_ = b.n_bytes;
    return b.n_records > 0;
}

}  // namespace mkh
