// Host feeder/consumer around the C ABI (include/merkurio_cuda.h): packs records into the engine's
// pinned slots, keeps several batches in flight on one or more GPUs, and hands the results back
// record by record, in input order — what the reference's inner loops (src/cmd_extract.rs:321-406,
// 463-607; src/cmd_tag.rs:530-557,585-612) do with one matcher call per record.
#pragma once
#include <deque>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "merkurio_cuda.h"

namespace mkh {

// What the command keeps per input record until its result arrives.
struct RecMeta {
    std::string a, b, c;  // extract: id, raw sequence text, quality — tag: name, SAM line
    uint32_t len = 0;     // bases
    uint8_t file = 0;     // 0 / 1: mate
    bool fastq = false, crlf = false;
    // chunked ingest (fastq_stream.h): the record is span `idx` of `chunk` and a, b, c stay empty; the
    // chunk is kept alive through Scanner::hold until the record has been delivered
    const void* chunk = nullptr;
    uint32_t idx = 0;
};

struct RecHit {
    uint64_t start;    // zero-based, in the record
    uint32_t pattern;  // index in the sorted pattern list
    uint32_t len;
};

// Called once per input record, in input order. `hits` is in Aho-Corasick report order
// (end asc, start asc, pattern asc); in MK_MODE_PATTERN_SET it holds one entry per distinct pattern
// (start = 0); in MK_MODE_FLAG it is empty and only `found` is meaningful.
using RecordCallback = std::function<void(RecMeta& meta, bool found, std::vector<RecHit>& hits)>;

class Scanner {
public:
    Scanner(const std::vector<std::string>& patterns, bool case_insensitive, mk_encoding enc, mk_mode mode, RecordCallback cb);
    ~Scanner();
    // ASCII sequence of one record (any length: long records are cut into overlapping pieces)
    void add_record(const char* seq, size_t len, RecMeta&& meta);
    // BAM 4-bit sequence of one record
    void add_record_packed(const uint8_t* packed, uint32_t l_seq, RecMeta&& meta);
    // Keep `owner` (the buffer the next records' metadata points into) alive until every record added
    // from now on has been delivered, plus one more batch (a pair's first mate may be delivered in the
    // batch before its second mate). One owner per input file (`file` = 0 / 1); a new call replaces it.
    void hold(int file, std::shared_ptr<const void> owner);
    void finish();  // flush the open batch and deliver every outstanding record
    int n_gpus() const { return (int)engines_.size(); }

private:
    struct Piece {
        bool first, last;
        uint64_t base;       // offset of the piece inside its record
        uint32_t own_from;   // hits ending at or before this piece-relative position belong to the previous piece
    };
    struct Batch {
        int engine = 0;
        uint32_t slot = 0;
        uint8_t* seq = nullptr;
        uint64_t* off = nullptr;
        uint32_t* lens = nullptr;
        uint32_t n_records = 0;
        uint64_t n_units = 0, n_bytes = 0;
        std::vector<Piece> pieces;
        std::vector<RecMeta> metas;  // one per piece with first == true
        std::vector<std::shared_ptr<const void>> owners;
    };
    void open_batch();
    void submit_open();
    void consume_oldest();
    void deliver_piece(Batch& b, uint32_t r, bool flag, const mk_hit* hits, size_t n);

    std::vector<mk_engine*> engines_;
    mk_encoding enc_;
    mk_mode mode_;
    RecordCallback cb_;
    uint32_t n_slots_ = 3, max_records_ = 0, max_pattern_len_ = 0;
    uint64_t max_bytes_ = 0, batch_seq_ = 0;
    std::unique_ptr<Batch> open_;
    std::deque<std::unique_ptr<Batch>> inflight_;
    std::vector<std::unique_ptr<Batch>> spare_;
    // record being assembled from its pieces
    RecMeta cur_meta_;
    bool cur_found_ = false;
    std::vector<RecHit> cur_hits_;
    double t_start_ = 0, t_setup_ = 0, t_wait_ = 0, t_consume_ = 0;  // MERKURIO_TIMING=1 prints them
    uint64_t device_ns_ = 0, n_records_ = 0, n_bases_ = 0;
    std::shared_ptr<const void> held_[2];
    std::vector<std::shared_ptr<const void>> grace_;  // owners of the batch consumed last
};

}  // namespace mkh
