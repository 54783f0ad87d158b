// Chunked SAM / BAM ingest for the tag feeder: one thread pulls the (inflated) bytes that follow the
// header into large blocks (block_reader.h), a second one indexes the alignment records in place; the packer thread then
// copies (BAM) or packs (SAM text) the sequences into the pinned batches. Replaces the per-record
// `bam` crate readers of src/cmd_tag.rs:504-531,561-586. Records are kept as they are in the file: a
// SAM line, or a BAM record body that is turned into SAM text only if the record is written.
#pragma once
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "block_reader.h"
#include "io.h"

namespace mkh {

struct AlnSpan {
    uint32_t off, len;            // SAM: the line without its break; BAM: the record body (after block_size)
    uint32_t name_off, name_len;
    uint32_t seq_off, l_seq;      // SAM: text (l_seq == 0 for "*"); BAM: 4-bit packed, (l_seq + 1) / 2 bytes
};

struct AlnChunk {
    ByteBuf data;
    std::vector<AlnSpan> recs;
    bool bam = false;
    std::string error;  // non-empty: the input is malformed right after recs.back()
};

class AlnChunkReader {
public:
    // Takes over the reader after its header has been parsed.
    explicit AlnChunkReader(std::unique_ptr<AlnReader> reader, size_t chunk_bytes = 8u << 20, size_t depth = 16);
    ~AlnChunkReader();
    std::shared_ptr<AlnChunk> next();  // nullptr after the last chunk
    const std::vector<std::string>& refs() const { return reader_->refs(); }
    bool is_bam() const { return reader_->is_bam(); }

private:
    struct Shared;
    static constexpr size_t kHead = 64u << 10;  // room in front of a block for the record its predecessor left unfinished
    void run();  // the indexing thread; the reading thread is blocks_'s
    std::unique_ptr<AlnReader> reader_;
    size_t chunk_bytes_, depth_;
    std::shared_ptr<Shared> pool_;
    std::unique_ptr<BlockReader> blocks_;
    std::thread thread_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<AlnChunk>> ready_;
    bool done_ = false, stop_ = false;
    std::string io_error_;
};

}  // namespace mkh
