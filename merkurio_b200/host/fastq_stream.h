// Chunked FASTQ ingest for the extract feeder: per input file one thread reads (or inflates) the
// file into large buffers and a second one indexes the 4-line records in place (line breaks are
// located 32 bytes at a time); the packer thread of the pipeline
// (fastq_pipeline.h) copies the sequence bytes into the pinned batch and the batch keeps (chunk,
// record) references for the writer — no per-record allocation. Replaces needletail's
// parse_fastx_file + per-record borrow in src/cmd_extract.rs:281,321-327,412,463-475 for FASTQ input;
// FASTA and anything unusual stays on FastxReader (io.h). Results are identical to FastxReader's by
// construction of the span rules below (tests/test_cli_cpu.py and tests/test_gpu_cli.py compare both).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "block_reader.h"
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

struct Chunk {
    ByteBuf data;
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
    static constexpr size_t kHead = 64u << 10;      // room in front of a block for the record its predecessor left unfinished
    static constexpr size_t kStretch = 128u << 10;  // bytes indexed at a time (stays in the core's cache)
    void run();  // the indexing thread; the reading thread is blocks_'s
    std::string path_;
    size_t chunk_bytes_, depth_;
    std::shared_ptr<Shared> pool_;
    std::unique_ptr<BlockReader> blocks_;
    std::thread thread_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Chunk>> ready_;
    bool done_ = false, stop_ = false;
    std::string io_error_;
    double t_index_ = 0, t_starved_ = 0, t_blocked_ = 0;  // MERKURIO_TIMING
};

}  // namespace mkh
