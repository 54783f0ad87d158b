#include "block_reader.h"

#include <chrono>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include "codecs.h"

namespace mkh {

void* big_alloc(size_t bytes) {
    const size_t align = (size_t)2 << 20;
    const size_t rounded = (bytes + align - 1) / align * align;
    void* p = std::aligned_alloc(align, rounded);
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    if (p) madvise(p, rounded, MADV_HUGEPAGE);  // advice only: failure changes nothing
#endif
    return p;
}

double steady_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

size_t prefetch_depth(size_t chunk_bytes, int n_files) {
    size_t budget = (size_t)1024 << 20;
    if (const char* mb = std::getenv("MERKURIO_PREFETCH_MB")) budget = (size_t)std::strtoull(mb, nullptr, 10) << 20;
    const size_t per_chunk = std::max<size_t>(chunk_bytes, 4096) + ((size_t)64 << 10);  // block + head room
    return std::min<size_t>(std::max<size_t>(budget / (size_t)std::max(n_files, 1) / per_chunk, 4), 4096);
}

void find_line_breaks(const char* d, size_t from, size_t to, OffsetList& out);

BlockReader::BlockReader(const std::string& path, size_t block_bytes, size_t head, size_t depth, bool scan_lines)
    : path_(path), block_bytes_(std::max<size_t>(block_bytes, 4096)), head_(head), depth_(std::max<size_t>(depth, 1)), scan_lines_(scan_lines) {
    helpers_ = std::getenv("MERKURIO_READ_THREADS") ? std::max(1, std::atoi(std::getenv("MERKURIO_READ_THREADS")))
                                                    : (int)std::max(1u, std::min(6u, std::thread::hardware_concurrency() * 3 / 8));
    // open here so that a missing file fails in the caller's thread, with the caller's context
    { std::unique_ptr<InputStream> probe = InputStream::open(path_); }
    thread_ = std::thread([this] { run(); });
}

BlockReader::BlockReader(ReadFn read, size_t block_bytes, size_t head, size_t depth)
    : read_(std::move(read)), block_bytes_(std::max<size_t>(block_bytes, 4096)), head_(head), depth_(std::max<size_t>(depth, 1)) {
    thread_ = std::thread([this] { run(); });
}

BlockReader::~BlockReader() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
}

bool BlockReader::next(Block& b) {
    std::unique_lock<std::mutex> lk(mu_);
    if (b.data.capacity() && spare_.size() < depth_ + 2) spare_.push_back(std::move(b.data));
    b.data = ByteBuf();
    b.n = 0;
    b.last = false;
    cv_.wait(lk, [this] { return !ready_.empty() || done_; });
    if (!ready_.empty()) {
        b = std::move(ready_.front());
        ready_.pop_front();
        lk.unlock();
        cv_.notify_all();
        return true;
    }
    if (!io_error_.empty()) throw Error(io_error_);
    return false;
}

void BlockReader::run() {
    std::string error;
    try {
        std::unique_ptr<InputStream> src;
        if (!read_) {
            src = InputStream::open(path_);
            InputStream* in = src.get();
            const int rt = helpers_;  // concurrent preads of one block of an uncompressed regular file
            read_ = [in, rt](char* dst, size_t n) { return in->read_parallel(dst, n, rt); };
        }
        for (bool eof = false; !eof;) {
            Block b;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!spare_.empty()) { b.data = std::move(spare_.back()); spare_.pop_back(); }
            }
            if (b.data.size() < head_ + block_bytes_) b.data.resize(head_ + block_bytes_);
            const double t0 = steady_seconds();
            try {
                while (!eof && b.n < block_bytes_) {
                    size_t got = read_(b.data.data() + head_ + b.n, block_bytes_ - b.n);
                    if (got == 0) eof = true;
                    b.n += got;
                }
            } catch (const std::exception& e) {
                // what was read before the error still goes out (as a block that is not the last one: its
                // unfinished record is not a truncation), the error after it
                error = e.what();
                if (b.n == 0) break;
            }
            b.last = eof && error.empty();
            b.nl.clear();
            b.nl_ctx.clear();
            b.has_nl = false;
            if (scan_lines_ && b.n > 0 && head_ + b.n < ((size_t)1 << 32)) {
                // line breaks of the block, a slice per helper thread, concatenated in order
                const int K = (b.n >= ((size_t)1 << 20)) ? helpers_ : 1;
                const char* d = b.data.data();
                const size_t lo = head_, hi = head_ + b.n;
                // the neighbours of the breaks list[i0...) (the whole block has been read: a slice may look beyond its end)
                auto context = [d, lo, hi](const OffsetList& list, size_t i0, uint8_t* out) {
                    for (size_t i = i0; i < list.n; ++i) {
                        const size_t e = list.p[i];
                        uint8_t c = 0;
                        if (e > lo) c |= d[e - 1] == '\r' ? kNlCr : 0;
                        else c |= kNlNoPrev;
                        if (e + 1 < hi) c |= d[e + 1] == '@' ? kNlAt : (d[e + 1] == '+' ? kNlPlus : (d[e + 1] == '>' ? kNlGt : 0));
                        else c |= kNlNoNext;
                        out[i - i0] = c;
                    }
                };
                if (K == 1) {
                    // a stretch at a time, so that the neighbours are looked at while the stretch is in the cache
                    for (size_t at = lo; at < hi;) {
                        const size_t upto = std::min(hi, at + ((size_t)128 << 10)), i0 = b.nl.n;
                        find_line_breaks(d, at, upto, b.nl);
                        b.nl_ctx.resize(b.nl.n);
                        context(b.nl, i0, b.nl_ctx.data() + i0);
                        at = upto;
                    }
                } else {
                    std::vector<OffsetList> part((size_t)K);
                    std::vector<std::vector<uint8_t>> part_ctx((size_t)K);
                    auto job = [&](int t) {
                        OffsetList& list = part[(size_t)t];
                        std::vector<uint8_t>& cx = part_ctx[(size_t)t];
                        const size_t from = lo + b.n * (size_t)t / (size_t)K, to = lo + b.n * (size_t)(t + 1) / (size_t)K;
                        for (size_t at = from; at < to;) {
                            const size_t upto = std::min(to, at + ((size_t)128 << 10)), i0 = list.n;
                            find_line_breaks(d, at, upto, list);
                            cx.resize(list.n);
                            context(list, i0, cx.data() + i0);
                            at = upto;
                        }
                    };
                    std::vector<std::thread> th;
                    for (int t = 1; t < K; ++t) th.emplace_back(job, t);
                    job(0);
                    for (auto& x : th) x.join();
                    for (auto& pl : part) b.nl.append(pl);
                    for (auto& cx : part_ctx) b.nl_ctx.insert(b.nl_ctx.end(), cx.begin(), cx.end());
                }
                b.has_nl = true;
            }
            const double dt = steady_seconds() - t0;
            std::unique_lock<std::mutex> lk(mu_);
            t_read_ += dt;
            cv_.wait(lk, [this] { return ready_.size() < depth_ || stop_; });
            if (stop_) return;
            ready_.push_back(std::move(b));
            lk.unlock();
            cv_.notify_all();
            if (!error.empty()) break;
        }
    } catch (const std::exception& e) {
        error = e.what();
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        io_error_ = error;
        done_ = true;
    }
    cv_.notify_all();
}

namespace {
#if defined(__x86_64__)
// TWO: also match the second byte
template <bool TWO>
__attribute__((target("avx2"))) void scan_avx2(const char* d, size_t from, size_t to, char c1, char c2, OffsetList& out) {
    const __m256i v1 = _mm256_set1_epi8(c1), v2 = _mm256_set1_epi8(c2);
    size_t p = from;
    while (p + 32 <= to) {
        const size_t stop = std::min(to, p + 4096) - 31;  // a stretch of at most 4 KiB: room for its matches up front
        out.reserve(out.n + 4096);
        uint32_t* w = out.p + out.n;
        for (; p < stop; p += 32) {
            const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(d + p));
            __m256i eq = _mm256_cmpeq_epi8(x, v1);
            if (TWO) eq = _mm256_or_si256(eq, _mm256_cmpeq_epi8(x, v2));
            uint32_t m = (uint32_t)_mm256_movemask_epi8(eq);
            while (m) {
                *w++ = (uint32_t)(p + (unsigned)__builtin_ctz(m));
                m &= m - 1;
            }
        }
        out.n = (size_t)(w - out.p);
    }
    out.reserve(out.n + 32);
    for (; p < to; ++p)
        if (d[p] == c1 || (TWO && d[p] == c2)) out.p[out.n++] = (uint32_t)p;
}
#endif

template <bool TWO>
void scan_bytes(const char* d, size_t from, size_t to, char c1, char c2, OffsetList& out) {
    size_t p = from;
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !std::getenv("MERKURIO_NO_AVX2");  // (the switch: tests of the SSE2 path)
    if (have_avx2) return scan_avx2<TWO>(d, from, to, c1, c2, out);
    const __m128i v1 = _mm_set1_epi8(c1), v2 = _mm_set1_epi8(c2);
    while (p + 16 <= to) {
        const size_t stop = std::min(to, p + 4096) - 15;
        out.reserve(out.n + 4096);
        uint32_t* w = out.p + out.n;
        for (; p < stop; p += 16) {
            const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(d + p));
            __m128i eq = _mm_cmpeq_epi8(x, v1);
            if (TWO) eq = _mm_or_si128(eq, _mm_cmpeq_epi8(x, v2));
            uint32_t m = (uint32_t)_mm_movemask_epi8(eq);
            while (m) {
                *w++ = (uint32_t)(p + (unsigned)__builtin_ctz(m));
                m &= m - 1;
            }
        }
        out.n = (size_t)(w - out.p);
    }
#endif
    for (; p < to; ++p)
        if (d[p] == c1 || (TWO && d[p] == c2)) {
            out.reserve(out.n + 1);
            out.p[out.n++] = (uint32_t)p;
        }
}
}  // namespace

void find_line_breaks(const char* d, size_t from, size_t to, OffsetList& out) { scan_bytes<false>(d, from, to, '\n', '\n', out); }
void find_breaks_and_tabs(const char* d, size_t from, size_t to, OffsetList& out) { scan_bytes<true>(d, from, to, '\n', '\t', out); }

}  // namespace mkh
