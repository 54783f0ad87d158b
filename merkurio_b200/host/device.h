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

// Set by main() for the one-command-per-process case: ~EngineSet leaves the engines (device and pinned
// memory, streams, context) to the end of the process instead of releasing them one by one.
extern bool g_leave_engines_to_process_exit;
void report_process_time_if_asked();  // MERKURIO_TIMING: the line exit() would have printed

// What the command keeps per input record until its result arrives.
struct RecMeta {
    std::string a, b, c;  // extract: id, raw sequence text, quality — tag: name, SAM line
    uint32_t len = 0;     // bases
    uint8_t file = 0;     // 0 / 1: mate
    bool fastq = false, crlf = false;
    // FASTQ pipeline (fastq_pipeline.h): the record is span `idx` of `chunk` and a, b, c stay empty; the
    // batch being delivered keeps the chunk alive
    const void* chunk = nullptr;
    std::shared_ptr<const void> keep;  // FASTA records: owner of `chunk` for a meta that outlives its batch (the first mate of a pair)
    uint32_t idx = 0;
    uint8_t kind = 0;  // 0: strings a / b / c; 1: FASTQ span (chunk = Chunk, idx); 2: FASTA record (chunk = FaRecord); 3: alignment span
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

// One engine per GPU (MERKURIO_GPUS, default 1), each with MERKURIO_SLOTS (default 3) pinned slots of
// MERKURIO_BATCH_MB (default 64) MiB. Batch i of a run goes to engine i % G, slot (i / G) % S.
// MERKURIO_TIMING=1 prints where the time went when the set is destroyed.
struct EngineSet {
    EngineSet(const std::vector<std::string>& patterns, bool case_insensitive, uint32_t default_batch_mb = 64);
    ~EngineSet();
    EngineSet(const EngineSet&) = delete;
    void wait(int engine, uint32_t slot, mk_result* out);  // mk_scan_wait + accounting
    std::vector<mk_engine*> engines;
    uint32_t n_slots = 3, max_records = 0, max_pattern_len = 0;
    uint64_t max_bytes = 0;
    double t_start = 0, t_setup = 0, t_wait = 0, t_deliver = 0;
    double t_run = 0;  // SlotPipeline::run from its first to its last statement
    double t_submit = 0, t_idle = 0;  // ... of which: inside mk_scan_submit / waiting for the packer with nothing in flight
    double t_pack = 0, t_pack_wait = 0;  // packer thread: filling slots / waiting for a free slot
    uint64_t device_ns = 0, n_records = 0, n_bases = 0, n_batches = 0;
};

class Scanner {
public:
    Scanner(EngineSet& engines, mk_encoding enc, mk_mode mode, RecordCallback cb);
    ~Scanner();
    // ASCII sequence of one record (any length: long records are cut into overlapping pieces)
    void add_record(const char* seq, size_t len, RecMeta&& meta);
    // BAM 4-bit sequence of one record
    void add_record_packed(const uint8_t* packed, uint32_t l_seq, RecMeta&& meta);
    void finish();  // flush the open batch and deliver every outstanding record
    int n_gpus() const { return (int)es_.engines.size(); }

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
    };
    void open_batch();
    void submit_open();
    void consume_oldest();
    void deliver_piece(Batch& b, uint32_t r, bool flag, const mk_hit* hits, size_t n);

    EngineSet& es_;
    std::vector<mk_engine*>& engines_;
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
};

}  // namespace mkh
