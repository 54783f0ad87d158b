// FASTQ -> GPU pipeline of `extract` (replaces the per-record loop of src/cmd_extract.rs:321-406 and
// :463-607 for FASTQ input). Three stages, each on its own thread(s):
//
//   reader  (one thread per input file, fastq_stream.h): read / inflate, index the records in place
//   packer  (one thread): copy the sequence bytes of successive records — mates interleaved for
//           paired input — straight into a pinned slot of an engine, until the slot is full
//   driver  (the calling thread): submit the slot, wait for the oldest batch in flight, hand
//           (batch, result) to the command's consumer, return the slot to the packer
//
// A batch keeps references to the chunks its records live in, so that the consumer can name and
// write the few records that matched without any per-record bookkeeping for those that did not.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "device.h"
#include "fastq_stream.h"

namespace mkh {

// `count` consecutive records of one chunk; batch records [rec0, rec0 + count) of that file.
struct BatchSeg {
    std::shared_ptr<Chunk> chunk;
    uint32_t first, count;
    uint32_t rec0;  // index among the batch's records of the same file
};

struct PackedBatch {
    int engine = 0;
    uint32_t slot = 0;
    uint8_t* seq = nullptr;
    uint64_t* off = nullptr;
    uint32_t n_records = 0;   // paired: 2 x pairs, record 2k = mate 1 of pair k, record 2k+1 = its mate 2
    uint64_t n_units = 0;
    std::vector<BatchSeg> seg[2];  // per input file
    // set on the last batch when the input ended with an error: raised after the batch is delivered
    std::vector<std::string> error_chain;
    // the record of file `f` with index i among the batch's records of that file
    const BatchSeg& locate(int f, uint32_t i, size_t* cursor) const {
        const std::vector<BatchSeg>& v = seg[f];
        size_t c = *cursor < v.size() ? *cursor : 0;
        if (i < v[c].rec0) c = 0;
        while (i >= v[c].rec0 + v[c].count) ++c;
        *cursor = c;
        return v[c];
    }
};

using BatchConsumer = std::function<void(const PackedBatch&, const mk_result&)>;

class FastqPipeline {
public:
    // The readers are created by the caller (before the engines, so that reading and indexing the input
    // overlaps CUDA start-up); reader2 == nullptr: single-end.
    FastqPipeline(EngineSet& engines, std::unique_ptr<FastqChunkReader> reader1, std::unique_ptr<FastqChunkReader> reader2, mk_mode mode,
                  BatchConsumer consumer);
    static std::unique_ptr<FastqChunkReader> open_reader(const std::string& path);
    ~FastqPipeline();
    // Runs the whole input. Throws the input's error (parse error, unequal files) after everything before
    // it has been delivered, like the reference, which fails at the record it cannot read.
    void run();

private:
    void pack();
    bool fill(PackedBatch& b);  // false: input exhausted and nothing was added
    EngineSet& es_;
    bool paired_;
    mk_mode mode_;
    BatchConsumer consumer_;
    std::thread packer_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::unique_ptr<PackedBatch>> free_, packed_;
    bool packer_done_ = false, stop_ = false;
    std::string packer_error_;
    // packer state
    std::unique_ptr<FastqChunkReader> rd_[2];
    std::shared_ptr<Chunk> cur_[2];
    size_t idx_[2] = {0, 0};
    bool input_done_ = false;
};

}  // namespace mkh
