#include "fasta_pipeline.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "codecs.h"

namespace mkh {

// ------------------------------------------------------------------------------------------------
struct FastaChunkReader::Shared {
    std::mutex mu;
    std::vector<std::unique_ptr<FaChunk>> free_list;
};

FastaChunkReader::FastaChunkReader(const std::string& path, size_t chunk_bytes, size_t depth)
    : path_(path), chunk_bytes_(std::max<size_t>(chunk_bytes, 4096)), depth_(std::max<size_t>(depth, 1)), pool_(new Shared) {
    blocks_.reset(new BlockReader(path_, chunk_bytes_, kHead, 3, true));  // a missing file fails here, in the caller's thread
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
    std::vector<char> carry;
    try {
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
            const size_t carry_len = carry.size();
            carry.clear();
            // line breaks are located a stretch at a time, the lines noted while the stretch is still in the cache
            const char* d = c->data.data();
            size_t p = begin;
            size_t scan_from = begin;
            if (rb.has_nl && carry_len <= kHead && rb.nl_ctx.size() == rb.nl.size()) {
                // the reading stage has located the block's line breaks and noted whether a '>' follows each
                // (BlockReader::kNlGt): the lines are listed without touching the block's bytes again. Only the bytes
                // carried over are scanned here.
                nl.clear();
                find_line_breaks(d, begin, kHead, nl);
                const size_t n0 = nl.n, n1 = rb.nl.size(), base = c->lines.size();
                c->lines.resize(base + n0 + n1);
                FaLine* w = c->lines.data() + base;
                for (size_t i = 0; i < n0; ++i) {
                    const size_t e = nl.p[i];
                    w[i] = FaLine{(uint32_t)p, (uint32_t)(e - p), (uint8_t)(e > p && d[p] == '>')};
                    p = e + 1;
                }
                w += n0;
                const uint8_t* cx = rb.nl_ctx.data();
                for (size_t i = 0; i < n1; ++i) {
                    const size_t e = rb.nl[i];
                    // the line starts behind the block's previous break (what follows it was noted), or in front of the block
                    const bool gt = (i > 0 && p == (size_t)rb.nl[i - 1] + 1) ? (cx[i - 1] & BlockReader::kNlGt) != 0 : (e > p && d[p] == '>');
                    w[i] = FaLine{(uint32_t)p, (uint32_t)(e - p), (uint8_t)gt};
                    p = e + 1;
                }
                scan_from = have;
            }
            for (size_t scanned = scan_from; scanned < have;) {
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
        // The input cannot be read any further. If the bytes in front of that spot end in an unfinished header line,
        // the record before it is complete — a line-by-line reader has seen the '>' — so that line is still handed
        // out; the error follows it.
        std::shared_ptr<FaChunk> tail;
        if (!carry.empty() && carry[0] == '>' && carry.size() < ((size_t)1 << 30)) {
            tail.reset(new FaChunk);
            ByteBuf d(kHead + carry.size());
            std::memcpy(d.data() + kHead, carry.data(), carry.size());
            tail->data.swap(d);
            tail->lines.push_back(FaLine{(uint32_t)kHead, (uint32_t)carry.size(), 1});
        }
        std::lock_guard<std::mutex> lk(mu_);
        if (tail && !stop_) ready_.push_back(std::move(tail));
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
namespace {
// n bytes, for the lengths FASTA lines have: 16-byte pieces, the last one overlapping its predecessor
inline void copy_line(uint8_t* dst, const char* src, size_t n) {
#if defined(__SSE2__)
    if (n >= 16 && n <= 256) {
        size_t i = 0;
        for (; i + 16 < n; i += 16) _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + n - 16), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + n - 16)));
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

const char* kParseError = "Error during FASTQ/A record parsing.";
const char* kFirstCtx = "Error during FASTQ record parsing of first file.";
const char* kSecondCtx = "Error during FASTQ record parsing of second file. Do the two input files contain the same number of records?";
const char* kUnequal = "The two input files have a different number of records. Please provide valid paired-end read files.";
}  // namespace

std::unique_ptr<FastaChunkReader> FastaPipeline::open_reader(const std::string& path, int n_files) {
    const size_t chunk_bytes = std::getenv("MERKURIO_CHUNK_BYTES") ? (size_t)std::strtoull(std::getenv("MERKURIO_CHUNK_BYTES"), nullptr, 10)
                                                                   : (size_t)8 << 20;
    return std::unique_ptr<FastaChunkReader>(new FastaChunkReader(path, chunk_bytes, prefetch_depth(chunk_bytes, n_files)));
}

FastaPipeline::FastaPipeline(EngineSet& engines, std::unique_ptr<FastaChunkReader> reader, std::unique_ptr<FastaChunkReader> reader2,
                             mk_mode mode, bool keep_text, BatchConsumer consumer)
    : SlotPipeline(engines, MK_ENC_ASCII, mode, std::move(consumer)), paired_((bool)reader2), keep_text_(keep_text) {
    src_[0].rd = std::move(reader);
    src_[1].rd = std::move(reader2);
    overlap_ = es_.max_pattern_len ? es_.max_pattern_len - 1 : 0;
}

FastaPipeline::~FastaPipeline() { stop_packer(); }

void FastaPipeline::begin() {
    for (int f = 0; f < (paired_ ? 2 : 1); ++f) {
        try {
            src_[f].cur = src_[f].rd->next();
        } catch (const Error& e) {  // surfaces in fill(), where it gets the record path's context
            src_[f].read_error = e.what();
        }
    }
}

void FastaPipeline::flush_range(Src& s) {
    if (s.range_open && s.rec && keep_text_) s.rec->raw.push_back(FaRecord::Range{s.cur, s.range_off, s.range_end - s.range_off});
    s.range_open = false;
}

// Fill one slot. One record is packed at a time — with two files the records of the first and the second file in
// turn (record 2i of the consumer's sequence = record i of file 1, record 2i + 1 = its mate) — and a record that does
// not fit goes on in the next batch, whichever file it is from.
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
    auto fail = [&](std::vector<std::string> chain) {
        b.error_chain = std::move(chain);
        input_done_ = true;
    };
    // step s to its next line; false if the file cannot be read any further (*why: what the record-by-record
    // reader would have thrown there). At the end of the file s.cur is null.
    auto advance = [&](Src& s, std::string* why) {
        if (!s.read_error.empty()) { *why = s.read_error; return false; }
        while (s.cur && s.line == s.cur->lines.size()) {
            flush_range(s);
            try {
                s.cur = s.rd->next();
            } catch (const Error& e) {
                s.cur = nullptr;
                s.read_error = e.what();
                *why = s.read_error;
                return false;
            }
            s.line = 0;
            s.line_pos = 0;
        }
        return true;
    };
    // open a piece of s.rec in this batch; false if the batch has no record slot left
    auto open_piece = [&](Src& s, bool first) {
        if (b.n_records >= max_rec) return false;
        uint64_t lead = first ? 0 : std::min<uint64_t>(overlap_, s.rec->len);
        if (lead) {  // the record's last bases again: they are the tail of the slot filled before this one
            std::memmove(b.seq + b.n_bytes, prev_seq_ + prev_bytes_ - lead, lead);
        }
        b.off[b.n_records] = b.n_bytes;
        info->pieces.push_back(FaPiece{s.rec, first, false, s.rec->len - lead, (uint32_t)lead});
        b.n_bytes += lead;
        b.n_records += 1;
        s.rec_open_piece = true;
        return true;
    };
    auto end_record = [&](Src& s) {  // s.rec's piece is the last one of this batch; the other file is next
        flush_range(s);
        info->pieces.back().last = true;
        s.rec.reset();
        s.rec_open_piece = false;
        if (paired_) turn_ ^= 1;
    };
    for (;;) {
        Src& s = src_[turn_];
        std::string why;
        if (!advance(s, &why)) {
            fail(paired_ ? std::vector<std::string>{turn_ == 0 ? kFirstCtx : kSecondCtx, why} : std::vector<std::string>{kParseError, why});
            break;
        }
        if (!s.cur) {  // end of this file
            if (s.rec) {
                if (!s.rec_open_piece && !open_piece(s, false)) break;
                end_record(s);
                if (paired_) continue;
            } else if (paired_) {
                if (turn_ == 1) {  // file 1 had one more record
                    fail({kSecondCtx});
                    break;
                }
                Src& o = src_[1];  // file 1 is exhausted: file 2 must be, too
                if (!advance(o, &why)) fail({why});
                else if (o.cur) fail({kUnequal});
            }
            input_done_ = true;
            break;
        }
        const FaLine& ln = s.cur->lines[s.line];
        const char* text = s.cur->data.data() + ln.off;
        const bool cr = ln.len && text[ln.len - 1] == '\r';
        if (!s.started) {
            if (ln.len == (cr ? 1u : 0u)) { ++s.line; continue; }  // blank lines before the first record
            s.started = true;
        }
        if (ln.header) {
            if (s.rec) {  // the previous record ends here
                if (!s.rec_open_piece && !open_piece(s, false)) break;
                end_record(s);
                continue;  // (with two files: to the other file's record; this header waits)
            }
            if (b.n_records >= max_rec) break;  // the header opens the next batch
            s.rec.reset(new FaRecord);
            s.rec->id.assign(text + 1, ln.len - 1 - (cr ? 1 : 0));
            s.rec->file = (uint8_t)turn_;
            s.rec->crlf = cr;
            open_piece(s, true);
            ++s.line;
            continue;
        }
        // a sequence line of s.rec
        if (!s.rec_open_piece) {
            if (b.n_bytes + std::min<uint64_t>(overlap_, s.rec->len) >= cap && b.n_records > 0) break;
            if (!open_piece(s, false)) break;
        }
        if (s.line_pos == 0) {
            // Whole lines that fit the slot, one after the other — nearly every line of a FASTA file: copied in a
            // tight loop (a libc memcpy call per 60-byte line was most of the packer's time). Whatever ends the run
            // — the chunk's last line, a header, a line that must be split — is left to the general code.
            const FaLine* L = s.cur->lines.data();
            const size_t n_lines = s.cur->lines.size();
            const char* base = s.cur->data.data();
            size_t i = s.line;
            uint64_t at = b.n_bytes;
            uint32_t run_end = 0;
            for (; i < n_lines && !L[i].header; ++i) {
                const char* t = base + L[i].off;
                uint32_t n = L[i].len;
                if (n && t[n - 1] == '\r') --n;
                if (at + n > cap) break;
                copy_line(b.seq + at, t, n);
                at += n;
                run_end = L[i].off + L[i].len;
            }
            if (i > s.line) {
                if (!s.range_open) { s.range_open = true; s.range_off = L[s.line].off; }
                s.range_end = run_end;
                s.rec->len += at - b.n_bytes;
                b.n_bytes = at;
                s.line = i;
                continue;
            }
            if (!s.range_open) { s.range_open = true; s.range_off = ln.off; }
            s.range_end = ln.off + ln.len;
        }
        const uint32_t bases = ln.len - (cr ? 1 : 0);
        const uint64_t room = cap - b.n_bytes;
        const uint64_t take = std::min<uint64_t>(bases - s.line_pos, room);
        if (take) {
            std::memcpy(b.seq + b.n_bytes, text + s.line_pos, take);
            b.n_bytes += take;
            s.rec->len += take;
            s.line_pos += (uint32_t)take;
        }
        if (s.line_pos == bases) {
            ++s.line;
            s.line_pos = 0;
            continue;
        }
        // the slot is full in the middle of the record: it goes on in the next batch
        s.rec_open_piece = false;
        break;
    }
    b.off[b.n_records] = b.n_bytes;
    b.n_units = b.total_bases = b.n_bytes;
    prev_seq_ = b.seq;
    prev_bytes_ = b.n_bytes;
    return b.n_records > 0 || !b.error_chain.empty();
}

}  // namespace mkh
