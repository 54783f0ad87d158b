#include "codecs.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cerrno>
#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "bgzf.h"
#include "common.h"
#include "inflate.h"
#include "io.h"

namespace mkh {

std::unique_ptr<InputStream> open_parallel_gzip(int fd);  // pgzip.cpp

namespace {

size_t read_fd(int fd, void* dst, size_t n) {
    for (;;) {
        ssize_t got = ::read(fd, dst, n);
        if (got < 0) {
            if (errno == EINTR) continue;
            throw Error(std::string("read failed: ") + std::strerror(errno));
        }
        return (size_t)got;
    }
}

class PlainStream : public InputStream {
public:
    explicit PlainStream(int fd) : fd_(fd) {
        struct stat st;
        regular_ = ::fstat(fd_, &st) == 0 && S_ISREG(st.st_mode);
    }
    ~PlainStream() override { ::close(fd_); }
    size_t read(char* dst, size_t n) override {
        size_t got = read_fd(fd_, dst, n);
        pos_ += got;
        return got;
    }
    size_t read_parallel(char* dst, size_t n, int threads) override {
        if (!regular_ || threads <= 1 || n < ((size_t)1 << 20)) return read(dst, n);
        // slices of the range [pos_, pos_ + n), each read with pread until it is full or the file ends
        const size_t slice = (n / (size_t)threads + 4095) & ~(size_t)4095;
        std::vector<size_t> got((size_t)threads, 0);
        std::vector<int> err((size_t)threads, 0);
        auto job = [&](int t) {
            const size_t lo = std::min(n, slice * (size_t)t), hi = (t == threads - 1) ? n : std::min(n, slice * (size_t)(t + 1));
            size_t done = 0;
            while (lo + done < hi) {
                ssize_t r = ::pread(fd_, dst + lo + done, hi - lo - done, (off_t)(pos_ + lo + done));
                if (r < 0 && errno == EINTR) continue;
                if (r < 0) { err[(size_t)t] = errno; break; }
                if (r == 0) break;  // end of the file
                done += (size_t)r;
            }
            got[(size_t)t] = done;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) th.emplace_back(job, t);
        job(0);
        for (auto& x : th) x.join();
        size_t total = 0;
        for (int t = 0; t < threads; ++t) {
            if (err[(size_t)t]) throw Error(std::string("read failed: ") + std::strerror(err[(size_t)t]));
            const size_t lo = std::min(n, slice * (size_t)t), hi = (t == threads - 1) ? n : std::min(n, slice * (size_t)(t + 1));
            total += got[(size_t)t];
            if (got[(size_t)t] < hi - lo) break;  // the file ended inside this slice: nothing valid follows
        }
        pos_ += total;
        // keep the descriptor's own offset in step (read() continues from it)
        ::lseek(fd_, (off_t)pos_, SEEK_SET);
        return total;
    }
private:
    int fd_;
    bool regular_ = false;
    uint64_t pos_ = 0;
};

// gzip (RFC 1952; also concatenated members, as `cat a.gz b.gz` and bgzip produce) through the decoder of inflate.h
// (MERKURIO_ZLIB_INFLATE=1: through zlib's inflate, the round-1 path, kept for comparison). The end of the file
// inside a member is an error wherever it falls, like in the bzip2 / xz / zstd streams below, and so is a member
// whose CRC-32 or length does not match its trailer. What was decoded in front of a damaged spot is handed out first.
class GzipStream : public InputStream {
public:
    explicit GzipStream(int fd) : fd_(fd), in_((1 << 20) + 64), win_(kHistory + kChunk + Inflater::kOutputMargin + 64) {}
    ~GzipStream() override { ::close(fd_); }
    size_t read(char* dst, size_t n) override {
        if (n == 0) return 0;
        while (pend_begin_ == pend_end_) {
            if (failed_) throw Error(error_);
            if (state_ == kFinished) return 0;
            try {
                decode_some();
            } catch (const Error& e) {
                failed_ = true;
                error_ = e.what();
            }
        }
        const size_t k = std::min(n, pend_end_ - pend_begin_);
        std::memcpy(dst, win_.data() + pend_begin_, k);
        pend_begin_ += k;
        return k;
    }

private:
    enum State { kHeader, kBody, kTrailer, kFinished };
    static constexpr size_t kHistory = 32u << 10, kChunk = 1u << 20;

    // more compressed bytes behind the unconsumed ones; false at the end of the file
    bool fill() {
        if (eof_) return false;
        if (in_pos_ > 0) {
            std::memmove(in_.data(), in_.data() + in_pos_, in_len_ - in_pos_);
            in_len_ -= in_pos_;
            in_pos_ = 0;
        }
        const size_t cap = in_.size() - 64;
        bool any = false;
        while (in_len_ < cap) {
            const size_t got = read_fd(fd_, in_.data() + in_len_, cap - in_len_);
            if (got == 0) { eof_ = true; break; }
            in_len_ += got;
            any = true;
        }
        return any;
    }
    size_t avail() const { return in_len_ - in_pos_; }

    // Parses the member header at in_pos_: 1 = done, 0 = more input needed.
    int parse_header() {
        size_t len = 0;
        const int rc = parse_gzip_header(in_.data() + in_pos_, avail(), &len);
        if (rc < 0) throw Error("Error while decompressing the input (gzip)");
        if (rc) in_pos_ += len;
        return rc;
    }

    // One step: a header, up to kChunk bytes of a member's data, or a trailer.
    void decode_some() {
        if (state_ == kHeader) {
            if (avail() == 0 && !fill()) {
                if (!first_member_) { state_ = kFinished; return; }  // clean end of the file behind a member
                throw Error("Error while decompressing the input (truncated gzip stream)");
            }
            while (!parse_header()) {
                if (avail() >= in_.size() - 64) throw Error("Error while decompressing the input (gzip)");  // a header of a megabyte
                if (!fill()) throw Error("Error while decompressing the input (truncated gzip stream)");
            }
            first_member_ = false;
            inf_.reset();
            crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
            isize_ = 0;
            hist_ = 0;
            out_ = 0;
            state_ = kBody;
            return;
        }
        if (state_ == kBody) {
            // slide: keep the last 32 KiB in front of the place the next chunk is decoded to
            if (out_ > kHistory) {
                std::memmove(win_.data(), win_.data() + out_ - kHistory, kHistory);
                out_ = kHistory;
                hist_ = kHistory;
            }
            const size_t begin = out_;
            uint8_t* const out_end = win_.data() + begin + kChunk + Inflater::kOutputMargin;
            for (;;) {
                const uint8_t* ip = in_.data() + in_pos_;
                uint8_t* op = win_.data() + out_;
                const Inflater::Status rc = inf_.run(&ip, in_.data() + in_len_, eof_, win_.data(), &op, out_end);
                in_pos_ = (size_t)(ip - in_.data());
                out_ = (size_t)(op - win_.data());
                if (rc == Inflater::kNeedInput) {
                    fill();  // (at the end of the file the next call runs with in_final)
                    continue;
                }
                if (rc == Inflater::kError) {
                    deliver(begin);
                    throw Error(eof_ && in_pos_ + 16 >= in_len_ ? "Error while decompressing the input (truncated gzip stream)"
                                                                 : "Error while decompressing the input (gzip)");
                }
                if (rc == Inflater::kStreamEnd) state_ = kTrailer;
                break;
            }
            deliver(begin);
            return;
        }
        if (state_ == kTrailer) {
            while (avail() < 8)
                if (!fill()) throw Error("Error while decompressing the input (truncated gzip stream)");
            const uint8_t* t = in_.data() + in_pos_;
            const uint32_t want_crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            const uint32_t want_len = t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
            if (want_crc != crc_ || want_len != isize_) throw Error("Error while decompressing the input (gzip)");
            in_pos_ += 8;
            state_ = kHeader;
        }
    }
    // [begin, out_) of the window was decoded by this step
    void deliver(size_t begin) {
        if (out_ == begin) return;
        crc_ = crc32_fast(crc_, win_.data() + begin, out_ - begin);
        isize_ += (uint32_t)(out_ - begin);
        pend_begin_ = begin;
        pend_end_ = out_;
    }

    int fd_;
    std::vector<uint8_t> in_, win_;
    size_t in_pos_ = 0, in_len_ = 0;
    bool eof_ = false, first_member_ = true, failed_ = false;
    std::string error_;
    State state_ = kHeader;
    Inflater inf_;
    size_t hist_ = 0, out_ = 0;  // window: [0, out_) holds the latest output of the member, at least its last 32 KiB
    size_t pend_begin_ = 0, pend_end_ = 0;  // decoded, not handed out yet
    uint32_t crc_ = 0, isize_ = 0;
};

// The same through zlib (MERKURIO_ZLIB_INFLATE=1). Not gzread: that function reports an incomplete stream only from
// gzclose() (Z_BUF_ERROR), so a file cut short would be taken for a complete, shorter input.
class ZlibGzipStream : public InputStream {
public:
    explicit ZlibGzipStream(int fd) : fd_(fd), in_(1 << 20) {
        std::memset(&z_, 0, sizeof z_);
        if (inflateInit2(&z_, 15 + 16) != Z_OK) { ::close(fd_); throw Error("cannot open the gzip stream"); }
    }
    ~ZlibGzipStream() override {
        inflateEnd(&z_);
        ::close(fd_);
    }
    size_t read(char* dst, size_t n) override {
        if (done_) return 0;
        // at most 256 KiB per call: a broken stream is reported for the whole call, so that much of the good data in
        // front of the damage is lost at most (the caller asks again for the rest)
        const unsigned want = (unsigned)std::min<size_t>(n, 256u << 10);
        z_.next_out = reinterpret_cast<Bytef*>(dst);
        z_.avail_out = want;
        while (z_.avail_out == want) {
            if (z_.avail_in == 0 && !eof_) {
                size_t got = read_fd(fd_, in_.data(), in_.size());
                if (got == 0) eof_ = true;
                z_.next_in = reinterpret_cast<Bytef*>(in_.data());
                z_.avail_in = (unsigned)got;
            }
            if (between_members_) {
                if (z_.avail_in == 0) { done_ = true; break; }  // clean end of the file
                if (inflateReset(&z_) != Z_OK) throw Error("Error while decompressing the input (gzip)");
                between_members_ = false;
            }
            if (z_.avail_in == 0 && eof_) throw Error("Error while decompressing the input (truncated gzip stream)");
            const int rc = inflate(&z_, Z_NO_FLUSH);
            if (rc == Z_STREAM_END) between_members_ = true;
            else if (rc != Z_OK && rc != Z_BUF_ERROR) throw Error("Error while decompressing the input (gzip)");
        }
        return want - z_.avail_out;
    }
private:
    int fd_;
    z_stream z_;
    std::vector<char> in_;
    bool eof_ = false, done_ = false, between_members_ = false;
};

class BgzfStream : public InputStream {
public:
    BgzfStream(int fd, int threads) : rd_(fd, threads) {}
    size_t read(char* dst, size_t n) override { return rd_.read(dst, n); }
private:
    BgzfReader rd_;
};

void* load_symbol(void* lib, const char* name, const char* libname) {
    void* p = dlsym(lib, name);
    if (!p) throw Error(std::string(libname) + " lacks " + name);
    return p;
}

// ---- bzip2 (libbz2.so.1.0) ---------------------------------------------------------------------
struct bz_stream {
    char* next_in; unsigned avail_in; unsigned total_in_lo32, total_in_hi32;
    char* next_out; unsigned avail_out; unsigned total_out_lo32, total_out_hi32;
    void* state; void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};

class Bzip2Stream : public InputStream {
public:
    explicit Bzip2Stream(int fd) : fd_(fd), in_(1 << 20) {
        lib_ = dlopen("libbz2.so.1.0", RTLD_NOW);
        if (!lib_) lib_ = dlopen("libbz2.so.1", RTLD_NOW);
        if (!lib_) { ::close(fd_); throw Error("bzip2 input needs libbz2.so.1.0, which is not installed"); }
        init_ = (int (*)(bz_stream*, int, int))load_symbol(lib_, "BZ2_bzDecompressInit", "libbz2");
        run_ = (int (*)(bz_stream*))load_symbol(lib_, "BZ2_bzDecompress", "libbz2");
        end_ = (int (*)(bz_stream*))load_symbol(lib_, "BZ2_bzDecompressEnd", "libbz2");
        std::memset(&s_, 0, sizeof s_);
        if (init_(&s_, 0, 0) != 0) throw Error("BZ2_bzDecompressInit failed");
        open_ = true;
    }
    ~Bzip2Stream() override {
        if (open_) end_(&s_);
        ::close(fd_);
    }
    size_t read(char* dst, size_t n) override {
        if (done_) return 0;
        s_.next_out = dst;
        s_.avail_out = (unsigned)std::min<size_t>(n, 1u << 30);
        const unsigned want = s_.avail_out;
        while (s_.avail_out == want) {
            if (s_.avail_in == 0 && !eof_) {
                size_t got = read_fd(fd_, in_.data(), in_.size());
                if (got == 0) eof_ = true;
                s_.next_in = in_.data();
                s_.avail_in = (unsigned)got;
            }
            if (!open_) {  // a further stream of a multi-stream file, or trailing nothing
                if (s_.avail_in == 0) { done_ = true; break; }
                char* ni = s_.next_in; unsigned ai = s_.avail_in;
                std::memset(&s_, 0, sizeof s_);
                if (init_(&s_, 0, 0) != 0) throw Error("BZ2_bzDecompressInit failed");
                open_ = true;
                s_.next_in = ni; s_.avail_in = ai;
                s_.next_out = dst + (want - want); s_.avail_out = want;
            }
            if (s_.avail_in == 0 && eof_) throw Error("Error while decompressing the input (truncated bzip2 stream)");
            int rc = run_(&s_);
            if (rc == 4) {  // BZ_STREAM_END
                end_(&s_);
                open_ = false;
                if (s_.avail_out != want) break;
            } else if (rc != 0) {
                throw Error("Error while decompressing the input (bzip2)");
            }
        }
        return want - s_.avail_out;
    }
private:
    int fd_;
    void* lib_ = nullptr;
    int (*init_)(bz_stream*, int, int) = nullptr;
    int (*run_)(bz_stream*) = nullptr;
    int (*end_)(bz_stream*) = nullptr;
    bz_stream s_;
    std::vector<char> in_;
    bool open_ = false, eof_ = false, done_ = false;
};

// ---- xz (liblzma.so.5) -------------------------------------------------------------------------
struct lzma_stream {
    const uint8_t* next_in; size_t avail_in; uint64_t total_in;
    uint8_t* next_out; size_t avail_out; uint64_t total_out;
    const void* allocator; void* internal;
    void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
    uint64_t reserved_int1, reserved_int2; size_t reserved_int3, reserved_int4;
    int reserved_enum1, reserved_enum2;
};

class XzStream : public InputStream {
public:
    explicit XzStream(int fd) : fd_(fd), in_(1 << 20) {
        lib_ = dlopen("liblzma.so.5", RTLD_NOW);
        if (!lib_) { ::close(fd_); throw Error("xz input needs liblzma.so.5, which is not installed"); }
        auto decoder = (int (*)(lzma_stream*, uint64_t, uint32_t))load_symbol(lib_, "lzma_stream_decoder", "liblzma");
        code_ = (int (*)(lzma_stream*, int))load_symbol(lib_, "lzma_code", "liblzma");
        end_ = (void (*)(lzma_stream*))load_symbol(lib_, "lzma_end", "liblzma");
        std::memset(&s_, 0, sizeof s_);
        if (decoder(&s_, UINT64_MAX, 0x08 /* LZMA_CONCATENATED */) != 0) throw Error("lzma_stream_decoder failed");
    }
    ~XzStream() override {
        end_(&s_);
        ::close(fd_);
    }
    size_t read(char* dst, size_t n) override {
        if (done_) return 0;
        s_.next_out = reinterpret_cast<uint8_t*>(dst);
        s_.avail_out = n;
        while (s_.avail_out == n) {
            if (s_.avail_in == 0 && !eof_) {
                size_t got = read_fd(fd_, in_.data(), in_.size());
                if (got == 0) eof_ = true;
                s_.next_in = reinterpret_cast<const uint8_t*>(in_.data());
                s_.avail_in = got;
            }
            int rc = code_(&s_, eof_ ? 3 /* LZMA_FINISH */ : 0 /* LZMA_RUN */);
            if (rc == 1) { done_ = true; break; }  // LZMA_STREAM_END
            if (rc != 0) throw Error("Error while decompressing the input (xz)");
        }
        return n - s_.avail_out;
    }
private:
    int fd_;
    void* lib_ = nullptr;
    int (*code_)(lzma_stream*, int) = nullptr;
    void (*end_)(lzma_stream*) = nullptr;
    lzma_stream s_;
    std::vector<char> in_;
    bool eof_ = false, done_ = false;
};

// ---- zstd (libzstd.so.1) -------------------------------------------------------------------------
// needletail's "compression" feature (the reference's Cargo.toml:26) reads Zstandard input as well
// (magic 28 B5 2F FD). Streaming API of libzstd, stable since 1.0: buffers are {pointer, size, position}.
struct zstd_in { const void* src; size_t size, pos; };
struct zstd_out { void* dst; size_t size, pos; };

class ZstdStream : public InputStream {
public:
    explicit ZstdStream(int fd) : fd_(fd), in_(1 << 20) {
        lib_ = dlopen("libzstd.so.1", RTLD_NOW);
        if (!lib_) { ::close(fd_); throw Error("zstd input needs libzstd.so.1, which is not installed"); }
        auto create = (void* (*)())load_symbol(lib_, "ZSTD_createDStream", "libzstd");
        auto init = (size_t (*)(void*))load_symbol(lib_, "ZSTD_initDStream", "libzstd");
        run_ = (size_t (*)(void*, zstd_out*, zstd_in*))load_symbol(lib_, "ZSTD_decompressStream", "libzstd");
        is_error_ = (unsigned (*)(size_t))load_symbol(lib_, "ZSTD_isError", "libzstd");
        free_ = (size_t (*)(void*))load_symbol(lib_, "ZSTD_freeDStream", "libzstd");
        ds_ = create();
        if (!ds_ || is_error_(init(ds_))) { ::close(fd_); throw Error("ZSTD_createDStream failed"); }
        buf_ = zstd_in{in_.data(), 0, 0};
    }
    ~ZstdStream() override {
        if (ds_) free_(ds_);
        ::close(fd_);
    }
    size_t read(char* dst, size_t n) override {
        zstd_out out{dst, n, 0};
        while (out.pos == 0) {
            if (buf_.pos == buf_.size && !eof_) {
                size_t got = read_fd(fd_, in_.data(), in_.size());
                if (got == 0) eof_ = true;
                buf_ = zstd_in{in_.data(), got, 0};
            }
            const bool no_input = buf_.pos == buf_.size;
            if (no_input && !mid_frame_) break;  // the end of the input, between two frames
            // (without input the call only hands out what the decoder still holds)
            const size_t rc = run_(ds_, &out, &buf_);  // 0: a frame is complete (more frames may follow)
            if (is_error_(rc)) throw Error("Error while decompressing the input (zstd)");
            mid_frame_ = rc != 0;
            if (no_input && out.pos == 0) throw Error("Error while decompressing the input (truncated zstd stream)");
        }
        return out.pos;
    }
private:
    int fd_;
    void* lib_ = nullptr;
    void* ds_ = nullptr;
    size_t (*run_)(void*, zstd_out*, zstd_in*) = nullptr;
    unsigned (*is_error_)(size_t) = nullptr;
    size_t (*free_)(void*) = nullptr;
    std::vector<char> in_;
    zstd_in buf_{nullptr, 0, 0};
    bool eof_ = false, mid_frame_ = false;
};

}  // namespace

std::unique_ptr<InputStream> InputStream::open(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error("No such file or directory (os error 2)");
    unsigned char m[6] = {0, 0, 0, 0, 0, 0};
    ssize_t n = ::pread(fd, m, sizeof m, 0);
    if (n >= 2 && m[0] == 0x1f && m[1] == 0x8b) {
        if (decompression_threads() > 1 && is_bgzf(fd)) return std::unique_ptr<InputStream>(new BgzfStream(fd, decompression_threads()));
        if (std::getenv("MERKURIO_ZLIB_INFLATE")) return std::unique_ptr<InputStream>(new ZlibGzipStream(fd));
        if (std::unique_ptr<InputStream> par = open_parallel_gzip(fd)) return par;  // large regular files: several threads (pgzip.cpp)
        return std::unique_ptr<InputStream>(new GzipStream(fd));
    }
    if (n >= 3 && m[0] == 'B' && m[1] == 'Z' && m[2] == 'h') return std::unique_ptr<InputStream>(new Bzip2Stream(fd));
    if (n >= 6 && m[0] == 0xFD && m[1] == '7' && m[2] == 'z' && m[3] == 'X' && m[4] == 'Z' && m[5] == 0x00)
        return std::unique_ptr<InputStream>(new XzStream(fd));
    if (n >= 4 && m[0] == 0x28 && m[1] == 0xB5 && m[2] == 0x2F && m[3] == 0xFD) return std::unique_ptr<InputStream>(new ZstdStream(fd));
    return std::unique_ptr<InputStream>(new PlainStream(fd));
}

}  // namespace mkh
