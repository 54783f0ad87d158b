// Chunked FASTQ ingest for the extract feeder: per input file one thread reads (or inflates) the
// file into large buffers and a second one indexes the 4-line records in place (line breaks are
// located 32 bytes at a time); the packer thread of the pipeline
// (fastq_pipeline.h) copies the sequence bytes into the pinned batch and the batch keeps (chunk,
// record) references for the writer — no per-record allocation. Replaces needletail's
// parse_fastx_file + per-record borrow in src/cmd_extract.rs:281,321-327,412,463-475 for FASTQ input;
// FASTA and anything unusual stays on FastxReader (io.h). Results are identical to FastxReader's by
// construction of the span rules below (tests/test_cli_cpu.py and tests/test_gpu_cli.py compare both).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <new>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

namespace mkh {

// One FASTQ record inside Chunk::data.
struct RecSpan {
    uint32_t start;    // the '@'
    uint32_t id_len;   // header line without '@' and without its line break
    uint32_t seq_off, seq_len;
    uint32_t qual_off;  // qual_len == seq_len (checked)
    uint32_t end;      // one past the record's last byte (its final '\n' if there is one)
    uint8_t crlf;      // header line ended in "\r\n": the writer then ends every line with "\r\n"
    uint8_t plain;     // bytes [start, end) are exactly "@id\nseq\n+\nqual\n": the writer copies them
};

// Growable array of offsets without value initialisation (the indexer appends through a raw pointer).
struct OffsetList {
    uint32_t* p = nullptr;
    size_t n = 0, cap = 0;
    OffsetList() = default;
    OffsetList(const OffsetList&) = delete;
    OffsetList& operator=(const OffsetList&) = delete;
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

struct Chunk {
    std::vector<char> data;
    size_t len = 0;
    std::vector<RecSpan> recs;
    OffsetList nl;  // offsets of the line breaks of data[0, len...), scratch of the indexer
    bool failed = false;  // malformed input right after recs.back(): the consumer raises the parse error there
    const char* id(const RecSpan& r) const { return data.data() + r.start + 1; }
    const char* seq(const RecSpan& r) const { return data.data() + r.seq_off; }
    const char* qual(const RecSpan& r) const { return data.data() + r.qual_off; }
};

// True if the file's first byte (after gzip decoding) is '@', i.e. needletail would parse it as FASTQ.
bool looks_like_fastq(const std::string& path);

class FastqChunkReader {
public:
    // chunk_bytes: size of the read buffers (a record larger than that grows its chunk)
    explicit FastqChunkReader(const std::string& path, size_t chunk_bytes = 16u << 20, size_t depth = 4);
    ~FastqChunkReader();
    FastqChunkReader(const FastqChunkReader&) = delete;
    // Next chunk in file order; nullptr after the last one. I/O errors are rethrown here.
    std::shared_ptr<Chunk> next();

private:
    struct Shared;  // free list of chunk buffers; outlives the reader while chunks are still referenced
    struct RawBlock {
        std::unique_ptr<Chunk> chunk;  // n bytes of the file at data[kHead, kHead + n)
        size_t n = 0;
        bool last = false;
    };
    static constexpr size_t kHead = 64u << 10;  // room in front of a block for the record its predecessor left unfinished
    static constexpr size_t kStretch = 128u << 10;  // bytes indexed at a time (stays in the core's cache)
    void read_blocks();  // thread 1: read / inflate
    void run();          // thread 2: index
    std::string path_;
    size_t chunk_bytes_, depth_;
    std::shared_ptr<Shared> pool_;
    std::thread io_thread_, thread_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<RawBlock> raw_;
    std::deque<std::shared_ptr<Chunk>> ready_;
    bool io_done_ = false, done_ = false, stop_ = false;
    std::string io_error_;
    double t_read_ = 0, t_index_ = 0, t_starved_ = 0, t_blocked_ = 0;  // MERKURIO_TIMING
};

}  // namespace mkh
