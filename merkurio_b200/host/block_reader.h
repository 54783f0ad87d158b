// First stage of the chunked readers (fastq_stream.h, fasta_pipeline.h): one thread that does nothing
// but read (or inflate, codecs.h) the input into large blocks, so that locating the records of a block
// (second stage, another thread) overlaps the read of the next one. Also the line-break scanner both
// indexers use.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

#include "common.h"

namespace mkh {

// Bytes that are not zero-filled when a buffer is sized (the reader overwrites them anyway; zero-filling
// 8 MB per new block cost as much as reading it from the page cache).
void* big_alloc(size_t bytes);  // 2 MB aligned, madvise(MADV_HUGEPAGE) where the system offers it; free() releases it

template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    DefaultInitAllocator() = default;
    template <class U> DefaultInitAllocator(const DefaultInitAllocator<U>&) {}
    // large buffers: 2 MB aligned and marked for transparent huge pages (512 x fewer page faults on first touch)
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes < kHugeFrom) return static_cast<T*>(::operator new(bytes));
        void* p = big_alloc(bytes);
        if (!p) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t n) {
        if (n * sizeof(T) < kHugeFrom) ::operator delete(p);
        else std::free(p);
    }
    static constexpr size_t kHugeFrom = (size_t)1 << 20;
    template <class U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using ByteBuf = std::vector<char, DefaultInitAllocator<char>>;

// Growable array of offsets without value initialisation (the scanner appends through a raw pointer).
struct OffsetList {
    uint32_t* p = nullptr;
    size_t n = 0, cap = 0;
    OffsetList() = default;
    OffsetList(const OffsetList&) = delete;
    OffsetList& operator=(const OffsetList&) = delete;
    OffsetList(OffsetList&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    OffsetList& operator=(OffsetList&& o) noexcept {
        if (this != &o) { std::free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; }
        return *this;
    }
    void append(const OffsetList& o) {
        if (!o.n) return;
        reserve(n + o.n);
        std::memcpy(p + n, o.p, o.n * sizeof(uint32_t));
        n += o.n;
    }
    ~OffsetList() { std::free(p); }
    void reserve(size_t want) {
        if (want <= cap) return;
        size_t c = std::max(want, cap * 2);
        void* q = std::realloc(p, c * sizeof(uint32_t));
        if (!q) throw std::bad_alloc();
        p = static_cast<uint32_t*>(q);
        cap = c;
    }
    void clear() { n = 0; }
    size_t size() const { return n; }
    uint32_t operator[](size_t i) const { return p[i]; }
};

class BlockReader {
public:
    // n bytes of the file at data[head, head + n); the head room is the indexer's (it puts the unfinished
    // record or line of the block before there).
    struct Block {
        ByteBuf data;
        size_t n = 0;
        bool last = false;  // the file ends with this block
        // scan_lines: the offsets (in data) of every '\n' of data[head, head + n), found by the reading stage on its
        // helper threads while the bytes are still in their caches
        OffsetList nl;
        bool has_nl = false;
        // and, per line break e, what stands either side of it (kNl* bits below), noted at the same time: the FASTQ
        // indexer then recognises ordinary records without touching the block's bytes again
        std::vector<uint8_t> nl_ctx;
    };
    static constexpr uint8_t kNlCr = 1;        // data[e - 1] == '\r'
    static constexpr uint8_t kNlAt = 2;        // data[e + 1] == '@'
    static constexpr uint8_t kNlPlus = 4;      // data[e + 1] == '+'
    static constexpr uint8_t kNlGt = 8;        // data[e + 1] == '>'
    static constexpr uint8_t kNlNoPrev = 64;   // e is the first byte of the block: bit kNlCr unknown
    static constexpr uint8_t kNlNoNext = 128;  // e is the last byte of the block: bits kNlAt / kNlPlus / kNlGt unknown
    BlockReader(const std::string& path, size_t block_bytes, size_t head, size_t depth = 3, bool scan_lines = false);
    // The same over an input that is already open: read(dst, n) returns up to n bytes, 0 at the end.
    using ReadFn = std::function<size_t(char*, size_t)>;
    BlockReader(ReadFn read, size_t block_bytes, size_t head, size_t depth = 3);
    ~BlockReader();
    BlockReader(const BlockReader&) = delete;
    // The next block in file order, swapped into b (whatever b.data held before is reused as a buffer).
    // False after the last block; an I/O error is rethrown here once the blocks before it are out.
    bool next(Block& b);
    size_t head() const { return head_; }
    size_t block_bytes() const { return block_bytes_; }
    double seconds_reading() {
        std::lock_guard<std::mutex> lk(mu_);
        return t_read_;
    }

private:
    void run();
    std::string path_;
    ReadFn read_;
    size_t block_bytes_, head_, depth_;
    bool scan_lines_ = false;
    int helpers_ = 1;  // MERKURIO_READ_THREADS: concurrent preads of one block (uncompressed regular files), line-break scans
    std::thread thread_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Block> ready_;
    std::vector<ByteBuf> spare_;
    bool done_ = false, stop_ = false;
    std::string io_error_;
    double t_read_ = 0;
};

// Offsets of every '\n' in d[from, to), appended to out in ascending order. One vector compare per 32
// (16) bytes instead of a memchr call per line: FASTA / FASTQ lines are short, so the per-call cost of
// memchr was most of the indexing time.
void find_line_breaks(const char* d, size_t from, size_t to, OffsetList& out);
// The same for every '\n' and every '\t' (SAM lines and their fields).
void find_breaks_and_tabs(const char* d, size_t from, size_t to, OffsetList& out);

double steady_seconds();

// How many chunks of chunk_bytes a chunked reader may hold ready: MERKURIO_PREFETCH_MB (default 1024) shared
// by the input files. The readers start before the engines, so this is what is read during CUDA start-up.
size_t prefetch_depth(size_t chunk_bytes, int n_files);

}  // namespace mkh
