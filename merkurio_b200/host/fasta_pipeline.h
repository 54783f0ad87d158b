// FASTA input of `extract` on the slot pipeline (slot_pipeline.h): replaces needletail's FASTA reader
// and the per-record loop of src/cmd_extract.rs:321-406 for (multi-line) FASTA. One thread reads 8 MB blocks
// (block_reader.h), a second one indexes their lines in place; the packer copies the sequence lines — line breaks
// dropped, exactly what record.seq() hands the reference's matchers — straight into the pinned slots.
// A record larger than what is left of a slot continues in the next batch as a further *piece* that
// starts with the last (longest pattern - 1) bases again, so that every occurrence lies inside one
// piece; a hit that ends inside that overlap belongs to the piece before. The record's wrapped text
// is never copied: the batch keeps (chunk, offset, length) ranges for the writer.
// Two FASTA files (paired reads, src/cmd_extract.rs:412-418, :463-607) go through the same packer: record i of the
// first file, then record i of the second, so the consumer sees the mates alternate in input order.
#pragma once
#include "fastq_stream.h"
#include "slot_pipeline.h"

namespace mkh {

struct FaLine {
    uint32_t off, len;  // without the line break; a trailing '\r' is still included in len
    uint8_t header;     // the line starts with '>'
};

struct FaChunk {
    ByteBuf data;
    std::vector<FaLine> lines;
    bool last = false;
};

class FastaChunkReader {
public:
    explicit FastaChunkReader(const std::string& path, size_t chunk_bytes = 8u << 20, size_t depth = 16);
    ~FastaChunkReader();
    std::shared_ptr<FaChunk> next();  // nullptr after the last chunk; I/O errors are rethrown here

private:
    struct Shared;
    static constexpr size_t kHead = 64u << 10;      // room in front of a block for the line its predecessor left unfinished
    static constexpr size_t kStretch = 128u << 10;  // bytes indexed at a time
    void run();  // the indexing thread; the reading thread is blocks_'s
    std::string path_;
    size_t chunk_bytes_, depth_;
    std::shared_ptr<Shared> pool_;
    std::unique_ptr<BlockReader> blocks_;
    std::thread thread_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<FaChunk>> ready_;
    bool done_ = false, stop_ = false;
    std::string io_error_;
};

// True if the first non-blank line of the (decompressed) file starts with '>'.
bool looks_like_fasta(const std::string& path);

// One FASTA record while its pieces travel through the batches.
struct FaRecord {
    std::string id;       // header without '>' and line break
    uint8_t file = 0;     // 0 / 1: which input file (mate)
    bool crlf = false;    // header ended in "\r\n"
    uint64_t len = 0;     // bases
    struct Range { std::shared_ptr<FaChunk> chunk; uint32_t off, len; };
    std::vector<Range> raw;  // the sequence lines as they are in the file (last line break excluded); empty if not kept
    // filled by the consumer
    bool found = false;
    std::vector<RecHit> hits;
};

struct FaPiece {
    std::shared_ptr<FaRecord> rec;
    bool first, last;
    uint64_t base;      // offset of the piece inside its record
    uint32_t own_from;  // hits ending at or before this piece-relative position belong to the previous piece
};

struct FaBatchInfo {
    std::vector<FaPiece> pieces;  // one per batch record
};

class FastaPipeline : public SlotPipeline {
public:
    // `reader2` (may be null): the second file of a pair. The input's errors (a file that cannot be read any
    // further, files with different numbers of records) are raised by run() after everything in front of them has
    // been delivered, with the record-by-record path's messages.
    FastaPipeline(EngineSet& engines, std::unique_ptr<FastaChunkReader> reader, std::unique_ptr<FastaChunkReader> reader2, mk_mode mode,
                  bool keep_text, BatchConsumer consumer);
    ~FastaPipeline() override;
    static std::unique_ptr<FastaChunkReader> open_reader(const std::string& path, int n_files);
    // the pieces of a batch handed to the consumer
    static const FaBatchInfo& info(const PackedBatch& b) { return *static_cast<const FaBatchInfo*>(b.extra.get()); }

protected:
    void begin() override;
    bool fill(PackedBatch& b) override;

private:
    // where the packer stands in one input file
    struct Src {
        std::unique_ptr<FastaChunkReader> rd;
        std::shared_ptr<FaChunk> cur;
        size_t line = 0;
        uint32_t line_pos = 0;  // bytes of the current sequence line already packed
        bool started = false;
        std::shared_ptr<FaRecord> rec;  // record being packed
        bool rec_open_piece = false;    // its current piece sits in the batch being filled
        // raw-text range of the record inside the current chunk
        bool range_open = false;
        uint32_t range_off = 0, range_end = 0;
        std::string read_error;  // what next() threw: the file ends there
    };
    void flush_range(Src& s);
    Src src_[2];
    bool paired_;
    int turn_ = 0;  // the file whose record is being packed
    bool keep_text_;
    uint64_t overlap_ = 0;
    const uint8_t* prev_seq_ = nullptr;  // the slot filled before this one (source of the overlap)
    uint64_t prev_bytes_ = 0;
};

}  // namespace mkh
