#include "io.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace mkh {

const char kNibbleChars[17] = "=ACMGRSVTWYHKDBN";

// ---------------------------------------------------------------------------------------------
static int g_decompression_threads = 0;
std::string spool_if_not_seekable(const std::string& path) {
    struct stat st;
    if (::stat(path.c_str(), &st) != 0 || S_ISREG(st.st_mode) || S_ISDIR(st.st_mode)) return path;
    const int in = ::open(path.c_str(), O_RDONLY);
    if (in < 0) return path;
    const char* dir = std::getenv("TMPDIR");
    std::string tmpl = std::string(dir && *dir ? dir : "/tmp") + "/merkurio-input-XXXXXX";
    const int out = ::mkstemp(&tmpl[0]);
    if (out < 0) { ::close(in); throw Error("cannot create a temporary file for the input stream " + path + ": " + std::strerror(errno)); }
    ::unlink(tmpl.c_str());
    std::vector<char> buf(4 << 20);
    for (;;) {
        ssize_t got = ::read(in, buf.data(), buf.size());
        if (got < 0 && errno == EINTR) continue;
        if (got < 0) { ::close(in); ::close(out); throw Error("read failed: " + std::string(std::strerror(errno))); }
        if (got == 0) break;
        for (ssize_t done = 0; done < got;) {
            ssize_t w = ::write(out, buf.data() + done, (size_t)(got - done));
            if (w < 0 && errno == EINTR) continue;
            if (w < 0) { ::close(in); ::close(out); throw Error("cannot spool the input stream " + path + ": " + std::strerror(errno)); }
            done += w;
        }
    }
    ::close(in);
    return "/proc/self/fd/" + std::to_string(out);  // `out` stays open for the rest of the process
}

void set_decompression_threads(int n) { g_decompression_threads = std::max(n, 1); }
int decompression_threads() {
    if (g_decompression_threads > 0) return g_decompression_threads;
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(8u, std::max(1u, hw / 2));
}

ByteSource::ByteSource(const std::string& path) : in_(InputStream::open(path)), buf_(1 << 22) {}
ByteSource::~ByteSource() {}

bool ByteSource::fill() {
    if (eof_) return false;
    if (pos_ < end_) std::memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
    end_ -= pos_;
    pos_ = 0;
    size_t n = in_->read(buf_.data() + end_, buf_.size() - end_);
    if (n == 0) { eof_ = true; return false; }
    end_ += n;
    return true;
}

int ByteSource::peek() {
    if (pos_ == end_ && !fill()) return -1;
    return (unsigned char)buf_[pos_];
}

bool ByteSource::getline(std::string* line) {
    line->clear();
    for (;;) {
        if (pos_ == end_ && !fill()) return !line->empty();
        const char* p = buf_.data() + pos_;
        const char* nl = (const char*)std::memchr(p, '\n', end_ - pos_);
        if (nl) {
            line->append(p, nl - p);
            pos_ += (size_t)(nl - p) + 1;
            return true;
        }
        line->append(p, end_ - pos_);
        pos_ = end_;
        if (eof_) return true;
    }
}

bool ByteSource::getline_view(const char** p, size_t* n, std::string* spill) {
    if (pos_ == end_ && !fill()) return false;
    const char* b = buf_.data() + pos_;
    const char* nl = (const char*)std::memchr(b, '\n', end_ - pos_);
    if (nl) {
        *p = b;
        *n = (size_t)(nl - b);
        pos_ += *n + 1;
        return true;
    }
    bool got = getline(spill);  // the line straddles the buffer end (or is the unterminated last one)
    *p = spill->data();
    *n = spill->size();
    return got;
}

size_t ByteSource::read_some(void* dst, size_t n) {
    if (pos_ == end_) {
        if (eof_) return 0;
        // nothing buffered: read straight into the caller's memory
        size_t got = in_->read(static_cast<char*>(dst), n);
        if (got == 0) eof_ = true;
        return got;
    }
    size_t take = std::min(n, end_ - pos_);
    std::memcpy(dst, buf_.data() + pos_, take);
    pos_ += take;
    return take;
}

bool ByteSource::read_exact(void* dst, size_t n) {
    char* d = (char*)dst;
    size_t got = 0;
    while (got < n) {
        if (pos_ == end_ && !fill()) {
            if (got == 0) return false;
            throw Error("unexpected end of file");
        }
        size_t take = std::min(n - got, end_ - pos_);
        std::memcpy(d + got, buf_.data() + pos_, take);
        pos_ += take;
        got += take;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
void FastxRecord::write(std::string* out) const {
    const char* le = crlf ? "\r\n" : "\n";
    if (!fastq) {
        *out += '>'; *out += id; *out += le; *out += raw; *out += le;
    } else {
        *out += '@'; *out += id; *out += le; *out += raw; *out += le; *out += '+'; *out += le; *out += qual; *out += le;
    }
}

FastxReader::FastxReader(const std::string& path) : src_(path) {}

static void strip_cr(std::string* s, bool* crlf) {
    if (!s->empty() && s->back() == '\r') { s->pop_back(); if (crlf) *crlf = true; }
}

bool FastxReader::next(FastxRecord* rec) {
    std::string line;
    if (!started_) {
        // skip leading blank lines, then decide the format from the first marker
        for (;;) {
            if (!src_.getline(&line)) return false;
            std::string t = line;
            strip_cr(&t, nullptr);
            if (!t.empty()) break;
        }
        started_ = true;
        if (line[0] == '>') fastq_ = false;
        else if (line[0] == '@') fastq_ = true;
        else throw Error("Error during FASTQ/A record parsing.");
        pending_ = line;
        have_pending_ = true;
    }
    if (!have_pending_) {
        if (!src_.getline(&line)) return false;
        if (fastq_) {  // tolerate trailing blank lines
            std::string t = line;
            strip_cr(&t, nullptr);
            while (t.empty()) {
                if (!src_.getline(&line)) return false;
                t = line;
                strip_cr(&t, nullptr);
            }
        }
        pending_ = line;
    }
    have_pending_ = false;
    rec->crlf = false;
    rec->fastq = fastq_;
    rec->seq.clear(); rec->raw.clear(); rec->qual.clear();
    std::string head = pending_;
    strip_cr(&head, &rec->crlf);
    if (fastq_) {
        if (head.empty() || head[0] != '@') throw Error("Error during FASTQ/A record parsing.");
        rec->id = head.substr(1);
        std::string plus;
        if (!src_.getline(&rec->seq) || !src_.getline(&plus) || !src_.getline(&rec->qual)) throw Error("Error during FASTQ/A record parsing.");
        strip_cr(&rec->seq, nullptr); strip_cr(&plus, nullptr); strip_cr(&rec->qual, nullptr);
        if (plus.empty() || plus[0] != '+' || rec->seq.size() != rec->qual.size()) throw Error("Error during FASTQ/A record parsing.");
        rec->raw = rec->seq;
        return true;
    }
    if (head.empty() || head[0] != '>') throw Error("Error during FASTQ/A record parsing.");
    rec->id = head.substr(1);
    bool first = true;
    for (;;) {
        int c = src_.peek();
        if (c < 0 || c == '>') break;
        const char* lp = nullptr;
        size_t ln = 0;
        src_.getline_view(&lp, &ln, &line);
        if (keep_raw_) {
            if (!first) rec->raw += '\n';
            rec->raw.append(lp, ln);  // keeps a '\r' of CRLF files inside the wrapped text, like the file
        }
        first = false;
        if (ln && lp[ln - 1] == '\r') --ln;
        rec->seq.append(lp, ln);
    }
    // the final line break of the record is not part of raw_seq
    if (!rec->raw.empty() && rec->raw.back() == '\r') rec->raw.pop_back();
    return true;
}

// ---------------------------------------------------------------------------------------------
struct PrefetchingFastxReader::State {
    FastxReader* rd;
    size_t budget, held = 0;  // bytes of record text the queue may hold / holds
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::unique_ptr<FastxRecord>> ready;
    std::unique_ptr<Error> error;  // raised after the records before it have been handed out
    bool done = false, stop = false;
};

static size_t record_bytes(const FastxRecord& r) { return r.id.size() + r.seq.size() + r.raw.size() + r.qual.size() + 64; }

PrefetchingFastxReader::PrefetchingFastxReader(FastxReader* reader, size_t budget_bytes) : st_(new State) {
    st_->rd = reader;
    st_->budget = budget_bytes;
    if (const char* mb = std::getenv("MERKURIO_PREFETCH_MB")) st_->budget = (size_t)std::strtoull(mb, nullptr, 10) << 20;
    State* s = st_.get();
    s->th = std::thread([s] {
        try {
            for (;;) {
                std::unique_ptr<FastxRecord> r(new FastxRecord);
                if (!s->rd->next(r.get())) break;
                const size_t bytes = record_bytes(*r);
                std::unique_lock<std::mutex> lk(s->mu);
                // at least one record is always allowed in, however large
                s->cv.wait(lk, [s, bytes] { return s->ready.empty() || s->held + bytes <= s->budget || s->stop; });
                if (s->stop) return;
                s->held += bytes;
                s->ready.push_back(std::move(r));
                lk.unlock();
                s->cv.notify_all();
            }
        } catch (const Error& e) {
            std::lock_guard<std::mutex> lk(s->mu);
            s->error.reset(new Error(e));
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(s->mu);
            s->error.reset(new Error(e.what()));
        }
        std::lock_guard<std::mutex> lk(s->mu);
        s->done = true;
        s->cv.notify_all();
    });
}

PrefetchingFastxReader::~PrefetchingFastxReader() {
    {
        std::lock_guard<std::mutex> lk(st_->mu);
        st_->stop = true;
    }
    st_->cv.notify_all();
    if (st_->th.joinable()) st_->th.join();
}

bool PrefetchingFastxReader::next(FastxRecord* rec) {
    std::unique_lock<std::mutex> lk(st_->mu);
    st_->cv.wait(lk, [this] { return !st_->ready.empty() || st_->done; });
    if (!st_->ready.empty()) {
        std::unique_ptr<FastxRecord> r = std::move(st_->ready.front());
        st_->ready.pop_front();
        st_->held -= std::min(st_->held, record_bytes(*r));
        lk.unlock();
        st_->cv.notify_all();
        std::swap(*rec, *r);
        return true;
    }
    if (st_->error) {
        Error e = *st_->error;
        st_->error.reset();
        throw e;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
uint8_t nibble_of_sam_char(char c) {
    if (c >= 'a' && c <= 'z') c = (char)(c - 0x20);
    const char* p = (const char*)std::memchr(kNibbleChars, c, 16);
    return p ? (uint8_t)(p - kNibbleChars) : 15;  // unknown characters are stored as N
}

AlnReader::AlnReader(const std::string& path, bool is_bam) : src_(path), bam_(is_bam) {
    if (bam_) {
        read_bam_header();
    } else {
        std::string line;
        while (src_.peek() == '@') {
            src_.getline(&line);
            if (!line.empty() && line.back() == '\r') line.pop_back();
            header_.push_back(line);
        }
    }
}

void AlnReader::read_bam_header() {
    char magic[4];
    int32_t l_text = 0, n_ref = 0;
    if (!src_.read_exact(magic, 4) || std::memcmp(magic, "BAM\1", 4) != 0) throw Error("not a BAM file (bad magic)");
    src_.read_exact(&l_text, 4);
    std::string text((size_t)l_text, '\0');
    if (l_text) src_.read_exact(&text[0], (size_t)l_text);
    while (!text.empty() && text.back() == '\0') text.pop_back();
    size_t pos = 0;
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        std::string ln = text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = nl == std::string::npos ? text.size() : nl + 1;
        if (!ln.empty()) header_.push_back(ln);
    }
    src_.read_exact(&n_ref, 4);
    for (int32_t i = 0; i < n_ref; ++i) {
        int32_t l_name = 0, l_ref = 0;
        src_.read_exact(&l_name, 4);
        std::string name((size_t)l_name, '\0');
        src_.read_exact(&name[0], (size_t)l_name);
        src_.read_exact(&l_ref, 4);
        if (!name.empty() && name.back() == '\0') name.pop_back();
        refs_.push_back(name);
    }
}

bool AlnReader::next(AlnRecord* rec) { return bam_ ? next_bam(rec) : next_sam(rec); }

bool AlnReader::next_sam(AlnRecord* rec) {
    std::string line;
    for (;;) {
        if (!src_.getline(&line)) return false;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty()) break;
    }
    // fields: QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL [tags]
    size_t starts[12];
    int nf = 0;
    starts[nf++] = 0;
    for (size_t i = 0; i < line.size() && nf < 12; ++i)
        if (line[i] == '\t') starts[nf++] = i + 1;
    if (nf < 11) throw Error("truncated record");
    rec->name = line.substr(0, starts[1] - 1);
    const char* seq = line.data() + starts[9];
    size_t slen = starts[10] - 1 - starts[9];
    if (slen == 1 && seq[0] == '*') slen = 0;
    rec->l_seq = (uint32_t)slen;
    rec->packed.assign((slen + 1) / 2, 0);
    for (size_t i = 0; i < slen; ++i) rec->packed[i >> 1] |= (uint8_t)(nibble_of_sam_char(seq[i]) << ((i & 1) ? 0 : 4));
    rec->sam_line = line;
    return true;
}

template <typename T>
static T rd(const char* b, size_t off) {
    T v;
    std::memcpy(&v, b + off, sizeof(T));
    return v;
}

static void append_num(std::string* s, long long v) { *s += std::to_string(v); }
static void append_float(std::string* s, float v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    s->append(buf, r.ptr);
}

void bam_body_to_sam(const char* b, size_t len, const std::vector<std::string>& refs, std::string* name, std::string* line,
                     const uint8_t** packed_out, uint32_t* l_seq_out) {
    if (len < 32) throw Error("truncated record");
    int32_t ref_id = rd<int32_t>(b, 0), pos = rd<int32_t>(b, 4);
    uint8_t l_read_name = rd<uint8_t>(b, 8), mapq = rd<uint8_t>(b, 9);
    uint16_t n_cigar = rd<uint16_t>(b, 12), flag = rd<uint16_t>(b, 14);
    int32_t l_seq = rd<int32_t>(b, 16), next_ref = rd<int32_t>(b, 20), next_pos = rd<int32_t>(b, 24), tlen = rd<int32_t>(b, 28);
    if (l_seq < 0 || 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq > len) throw Error("truncated record");
    if ((ref_id >= 0 && (size_t)ref_id >= refs.size()) || (next_ref >= 0 && (size_t)next_ref >= refs.size())) throw Error("reference id out of range");
    size_t q = 32;
    name->assign(b + q, l_read_name ? l_read_name - 1 : 0);
    q += l_read_name;
    std::string cigar;
    for (uint16_t i = 0; i < n_cigar; ++i) {
        uint32_t c = rd<uint32_t>(b, q + 4 * i);
        append_num(&cigar, c >> 4);
        cigar += "MIDNSHP=X"[c & 0xF];
    }
    if (cigar.empty()) cigar = "*";
    q += 4 * (size_t)n_cigar;
    size_t nbytes = ((size_t)l_seq + 1) / 2;
    const uint8_t* packed = reinterpret_cast<const uint8_t*>(b) + q;
    if (packed_out) *packed_out = packed;
    if (l_seq_out) *l_seq_out = (uint32_t)l_seq;
    q += nbytes;
    std::string seq((size_t)l_seq, '\0'), qual((size_t)l_seq, '\0');
    for (int32_t i = 0; i < l_seq; ++i) seq[(size_t)i] = kNibbleChars[(packed[(size_t)i >> 1] >> ((i & 1) ? 0 : 4)) & 0xF];
    bool no_qual = l_seq == 0 || (uint8_t)b[q] == 0xFF;
    for (int32_t i = 0; i < l_seq; ++i) qual[(size_t)i] = (char)((uint8_t)b[q + (size_t)i] + 33);
    q += (size_t)l_seq;
    std::string& s = *line;
    s = *name; s += '\t';
    append_num(&s, flag); s += '\t';
    s += ref_id >= 0 ? refs[(size_t)ref_id] : "*"; s += '\t';
    append_num(&s, (long long)pos + 1); s += '\t';
    append_num(&s, mapq); s += '\t';
    s += cigar; s += '\t';
    s += next_ref < 0 ? "*" : (next_ref == ref_id ? "=" : refs[(size_t)next_ref]); s += '\t';
    append_num(&s, (long long)next_pos + 1); s += '\t';
    append_num(&s, tlen); s += '\t';
    s += l_seq ? seq : "*"; s += '\t';
    s += no_qual ? "*" : qual;
    auto need = [&](size_t n) { if (q + n > len) throw Error("truncated record"); };
    while (q + 3 <= len) {  // optional fields
        s += '\t';
        s.append(b + q, 2);
        char typ = b[q + 2];
        q += 3;
        switch (typ) {
            case 'A': need(1); s += ":A:"; s += b[q]; q += 1; break;
            case 'c': need(1); s += ":i:"; append_num(&s, rd<int8_t>(b, q)); q += 1; break;
            case 'C': need(1); s += ":i:"; append_num(&s, rd<uint8_t>(b, q)); q += 1; break;
            case 's': need(2); s += ":i:"; append_num(&s, rd<int16_t>(b, q)); q += 2; break;
            case 'S': need(2); s += ":i:"; append_num(&s, rd<uint16_t>(b, q)); q += 2; break;
            case 'i': need(4); s += ":i:"; append_num(&s, rd<int32_t>(b, q)); q += 4; break;
            case 'I': need(4); s += ":i:"; append_num(&s, rd<uint32_t>(b, q)); q += 4; break;
            case 'f': need(4); s += ":f:"; append_float(&s, rd<float>(b, q)); q += 4; break;
            case 'Z': case 'H': {
                s += ':'; s += typ; s += ':';
                size_t e = q;
                while (e < len && b[e]) ++e;
                s.append(b + q, e - q);
                q = e + 1;
                break;
            }
            case 'B': {
                need(5);
                char sub = b[q];
                uint32_t n = rd<uint32_t>(b, q + 1);
                q += 5;
                s += ":B:"; s += sub;
                for (uint32_t i = 0; i < n; ++i) {
                    s += ',';
                    switch (sub) {
                        case 'c': need(1); append_num(&s, rd<int8_t>(b, q)); q += 1; break;
                        case 'C': need(1); append_num(&s, rd<uint8_t>(b, q)); q += 1; break;
                        case 's': need(2); append_num(&s, rd<int16_t>(b, q)); q += 2; break;
                        case 'S': need(2); append_num(&s, rd<uint16_t>(b, q)); q += 2; break;
                        case 'i': need(4); append_num(&s, rd<int32_t>(b, q)); q += 4; break;
                        case 'I': need(4); append_num(&s, rd<uint32_t>(b, q)); q += 4; break;
                        case 'f': need(4); append_float(&s, rd<float>(b, q)); q += 4; break;
                        default: throw Error("bad B-array subtype");
                    }
                }
                break;
            }
            default: throw Error("bad tag type");
        }
    }
}

bool AlnReader::next_bam(AlnRecord* rec) {
    int32_t block_size = 0;
    if (!src_.read_exact(&block_size, 4)) return false;
    if (block_size < 0) throw Error("truncated record");
    std::vector<char> b((size_t)block_size);
    if (block_size && !src_.read_exact(b.data(), b.size())) throw Error("unexpected end of file");
    const uint8_t* packed = nullptr;
    bam_body_to_sam(b.data(), b.size(), refs_, &rec->name, &rec->sam_line, &packed, &rec->l_seq);
    rec->packed.assign(packed, packed + (rec->l_seq + 1) / 2);
    return true;
}

}  // namespace mkh
