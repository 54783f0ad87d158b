#include "bamwriter.h"
#include "inflate.h"

#include <atomic>
#include <mutex>
#include <thread>

#include <zlib.h>

#include <cstdlib>
#include <cstring>

#include "io.h"

namespace mkh {

namespace {

constexpr size_t kBlockData = 0xff00;  // uncompressed bytes per BGZF block

template <typename T>
void push(std::vector<uint8_t>* v, T x) {
    uint8_t b[sizeof(T)];
    std::memcpy(b, &x, sizeof(T));
    v->insert(v->end(), b, b + sizeof(T));
}

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    size_t pos = 0;
    for (;;) {
        size_t t = s.find(sep, pos);
        out.push_back(s.substr(pos, t == std::string::npos ? std::string::npos : t - pos));
        if (t == std::string::npos) break;
        pos = t + 1;
    }
    return out;
}

// UCSC binning scheme (SAM spec 5.3)
uint16_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

void push_int_tag(std::vector<uint8_t>* v, long long x) {
    if (x >= 0) {
        if (x <= 0xFF) { v->push_back('C'); push<uint8_t>(v, (uint8_t)x); }
        else if (x <= 0xFFFF) { v->push_back('S'); push<uint16_t>(v, (uint16_t)x); }
        else { v->push_back('I'); push<uint32_t>(v, (uint32_t)x); }
    } else {
        if (x >= -128) { v->push_back('c'); push<int8_t>(v, (int8_t)x); }
        else if (x >= -32768) { v->push_back('s'); push<int16_t>(v, (int16_t)x); }
        else { v->push_back('i'); push<int32_t>(v, (int32_t)x); }
    }
}

}  // namespace

BamWriter::BamWriter(const std::string& path, const std::vector<std::string>& header_lines, int threads) : threads_(std::max(threads, 1)) {
    f_ = std::fopen(path.c_str(), "wb");
    if (!f_) throw Error("No such file or directory (os error 2)").with_context("Error writing BAM file: " + path);
    std::string text;
    std::vector<std::pair<std::string, int32_t>> refs;
    for (auto& ln : header_lines) {
        text += ln;
        text += '\n';
        if (ln.rfind("@SQ", 0) == 0) {
            std::string name;
            int32_t len = 0;
            for (auto& f : split(ln, '\t')) {
                if (f.rfind("SN:", 0) == 0) name = f.substr(3);
                if (f.rfind("LN:", 0) == 0) len = (int32_t)std::atoll(f.c_str() + 3);
            }
            ref_ids_[name] = (int32_t)refs.size();
            refs.emplace_back(name, len);
            ref_names_.push_back(name);
        }
    }
    std::vector<uint8_t> h;
    h.insert(h.end(), {'B', 'A', 'M', 1});
    push<int32_t>(&h, (int32_t)text.size());
    h.insert(h.end(), text.begin(), text.end());
    push<int32_t>(&h, (int32_t)refs.size());
    for (auto& r : refs) {
        push<int32_t>(&h, (int32_t)r.first.size() + 1);
        h.insert(h.end(), r.first.begin(), r.first.end());
        h.push_back(0);
        push<int32_t>(&h, r.second);
    }
    put(h.data(), h.size());
    flush_block();  // keep the header in its own block(s), like samtools
}

BamWriter::~BamWriter() {
    try { close(); } catch (...) {}
}

void BamWriter::put(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    while (n) {
        size_t take = std::min(n, kBlockData - buf_.size());
        buf_.insert(buf_.end(), b, b + take);
        b += take;
        n -= take;
        if (buf_.size() == kBlockData) flush_block();
    }
}

// One BGZF block: header with the block size, raw deflate of `in`, CRC32 and input size.
static void bgzf_compress_block(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {
    out->resize(0x10000 + 64);
    z_stream zs{};
    if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Error("deflateInit2 failed");
    zs.next_in = const_cast<uint8_t*>(in.data());
    zs.avail_in = (uInt)in.size();
    zs.next_out = out->data() + 18;
    zs.avail_out = (uInt)(out->size() - 18 - 8);
    int rc = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) throw Error("deflate failed while writing BAM");
    size_t clen = zs.total_out, total = 18 + clen + 8;
    const uint8_t head[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, (uint8_t)((total - 1) & 0xFF), (uint8_t)((total - 1) >> 8)};
    std::memcpy(out->data(), head, 18);
    uint32_t crc = crc32_fast(0, reinterpret_cast<const uint8_t*>(in.data()), in.size()), isize = (uint32_t)in.size();
    std::memcpy(out->data() + 18 + clen, &crc, 4);
    std::memcpy(out->data() + 18 + clen + 4, &isize, 4);
    out->resize(total);
}

void BamWriter::flush_block() {
    if (!f_ || buf_.empty()) return;
    pending_.emplace_back();
    pending_.back().swap(buf_);
    buf_.reserve(kBlockData);
    if (pending_.size() >= 64) compress_pending();
}

// Compress the waiting blocks on `threads_` threads and write them in order.
void BamWriter::compress_pending() {
    if (!f_ || pending_.empty()) return;
    std::vector<std::vector<uint8_t>> out(pending_.size());
    std::atomic<size_t> next{0};
    std::string error;
    std::mutex mu;
    auto work = [&] {
        try {
            for (size_t i; (i = next.fetch_add(1)) < pending_.size();) bgzf_compress_block(pending_[i], &out[i]);
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(mu);
            error = e.what();
        }
    };
    std::vector<std::thread> pool;
    const size_t extra = std::min<size_t>(threads_ > 1 ? (size_t)threads_ - 1 : 0, pending_.size() > 1 ? pending_.size() - 1 : 0);
    for (size_t t = 0; t < extra; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (!error.empty()) throw Error(error);
    for (auto& o : out) write_all(f_, o.data(), o.size());
    pending_.clear();
}

void BamWriter::close() {
    if (!f_) return;
    flush_block();
    compress_pending();
    static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    FILE* f = f_;
    f_ = nullptr;
    try {
        write_all(f, eof, sizeof eof);
    } catch (...) {
        std::fclose(f);
        throw;
    }
    close_checked(f);
}

void BamWriter::write_bam_record(const char* body, size_t len, const std::string& tag, const std::string& value) {
    int32_t block_size = (int32_t)(len + 3 + value.size() + 1);
    put(&block_size, 4);
    put(body, len);
    const char head[3] = {tag[0], tag[1], 'Z'};
    put(head, 3);
    put(value.c_str(), value.size() + 1);
}

int bam_find_tag(const char* b, size_t len, const std::string& tag, std::string* val) {
    if (len < 32) return 0;
    uint8_t l_read_name = (uint8_t)b[8];
    uint16_t n_cigar;
    int32_t l_seq;
    std::memcpy(&n_cigar, b + 12, 2);
    std::memcpy(&l_seq, b + 16, 4);
    size_t q = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    auto elem = [](char t) -> size_t { return (t == 'c' || t == 'C' || t == 'A') ? 1 : (t == 's' || t == 'S') ? 2 : (t == 'i' || t == 'I' || t == 'f') ? 4 : 0; };
    while (q + 3 <= len) {
        const bool hit = b[q] == tag[0] && b[q + 1] == tag[1];
        const char typ = b[q + 2];
        q += 3;
        if (typ == 'Z' || typ == 'H') {
            size_t e = q;
            while (e < len && b[e]) ++e;
            if (hit) {
                if (typ != 'Z') return 2;
                val->assign(b + q, e - q);
                return 1;
            }
            q = e + 1;
        } else if (typ == 'B') {
            if (hit) return 2;
            if (q + 5 > len) return 0;
            uint32_t n;
            std::memcpy(&n, b + q + 1, 4);
            size_t w = elem(b[q]);
            if (!w) return 0;
            q += 5 + (size_t)n * w;
        } else {
            if (hit) return 2;
            size_t w = elem(typ);
            if (!w) return 0;
            q += w;
        }
    }
    return 0;
}

void BamWriter::write_sam_line(const std::string& line) {
    std::vector<std::string> f = split(line, '\t');
    if (f.size() < 11) throw Error("truncated SAM record");
    auto ref_id = [&](const std::string& name) -> int32_t {
        if (name == "*") return -1;
        auto it = ref_ids_.find(name);
        if (it == ref_ids_.end()) throw Error("reference '" + name + "' is not in the header");
        return it->second;
    };
    const int32_t rid = ref_id(f[2]);
    const int32_t pos = (int32_t)std::atoll(f[3].c_str()) - 1;
    std::vector<uint32_t> cigar;
    int64_t ref_len = 0;
    if (f[5] != "*") {
        const char* p = f[5].c_str();
        while (*p) {
            char* e;
            unsigned long n = std::strtoul(p, &e, 10);
            const char* ops = "MIDNSHP=X";
            const char* o = std::strchr(ops, *e);
            if (!o || !*e) throw Error("bad CIGAR '" + f[5] + "'");
            uint32_t op = (uint32_t)(o - ops);
            cigar.push_back((uint32_t)(n << 4) | op);
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += n;
            p = e + 1;
        }
    }
    const std::string& seq = f[9];
    const uint32_t l_seq = seq == "*" ? 0 : (uint32_t)seq.size();
    std::vector<uint8_t> r;
    push<int32_t>(&r, rid);
    push<int32_t>(&r, pos);
    push<uint8_t>(&r, (uint8_t)(f[0].size() + 1));
    push<uint8_t>(&r, (uint8_t)std::atoi(f[4].c_str()));
    push<uint16_t>(&r, reg2bin(pos < 0 ? 0 : pos, (pos < 0 ? 0 : pos) + (ref_len > 0 ? ref_len : 1)));
    push<uint16_t>(&r, (uint16_t)cigar.size());
    push<uint16_t>(&r, (uint16_t)std::atoi(f[1].c_str()));
    push<int32_t>(&r, (int32_t)l_seq);
    push<int32_t>(&r, f[6] == "=" ? rid : ref_id(f[6]));
    push<int32_t>(&r, (int32_t)std::atoll(f[7].c_str()) - 1);
    push<int32_t>(&r, (int32_t)std::atoll(f[8].c_str()));
    r.insert(r.end(), f[0].begin(), f[0].end());
    r.push_back(0);
    for (uint32_t c : cigar) push<uint32_t>(&r, c);
    for (uint32_t i = 0; i < l_seq; i += 2) {
        auto nib = [&](char c) -> uint8_t {
            if (c >= 'a' && c <= 'z') c = (char)(c - 0x20);
            const char* q = (const char*)std::memchr(kNibbleChars, c, 16);
            return q ? (uint8_t)(q - kNibbleChars) : 15;
        };
        r.push_back((uint8_t)((nib(seq[i]) << 4) | (i + 1 < l_seq ? nib(seq[i + 1]) : 0)));
    }
    if (f[10] == "*") r.insert(r.end(), l_seq, 0xFF);
    else for (uint32_t i = 0; i < l_seq; ++i) r.push_back((uint8_t)(f[10][i] - 33));
    for (size_t k = 11; k < f.size(); ++k) {
        const std::string& t = f[k];
        if (t.size() < 5 || t[2] != ':' || t[4] != ':') throw Error("bad optional field '" + t + "'");
        r.push_back((uint8_t)t[0]);
        r.push_back((uint8_t)t[1]);
        const char* val = t.c_str() + 5;
        switch (t[3]) {
            case 'A': r.push_back('A'); r.push_back((uint8_t)val[0]); break;
            case 'i': push_int_tag(&r, std::atoll(val)); break;
            case 'f': { r.push_back('f'); push<float>(&r, std::strtof(val, nullptr)); break; }
            case 'Z': case 'H': r.push_back((uint8_t)t[3]); r.insert(r.end(), val, val + std::strlen(val)); r.push_back(0); break;
            case 'B': {
                std::vector<std::string> items = split(val, ',');
                char sub = items[0].empty() ? 'c' : items[0][0];
                r.push_back('B'); r.push_back((uint8_t)sub);
                push<uint32_t>(&r, (uint32_t)items.size() - 1);
                for (size_t i = 1; i < items.size(); ++i) {
                    switch (sub) {
                        case 'c': push<int8_t>(&r, (int8_t)std::atoi(items[i].c_str())); break;
                        case 'C': push<uint8_t>(&r, (uint8_t)std::atoi(items[i].c_str())); break;
                        case 's': push<int16_t>(&r, (int16_t)std::atoi(items[i].c_str())); break;
                        case 'S': push<uint16_t>(&r, (uint16_t)std::atoi(items[i].c_str())); break;
                        case 'i': push<int32_t>(&r, (int32_t)std::atoll(items[i].c_str())); break;
                        case 'I': push<uint32_t>(&r, (uint32_t)std::atoll(items[i].c_str())); break;
                        case 'f': push<float>(&r, std::strtof(items[i].c_str(), nullptr)); break;
                        default: throw Error("bad B-array subtype");
                    }
                }
                break;
            }
            default: throw Error("bad optional field type in '" + t + "'");
        }
    }
    int32_t block_size = (int32_t)r.size();
    put(&block_size, 4);
    put(r.data(), r.size());
}

}  // namespace mkh
