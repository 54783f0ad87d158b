// FASTQ input of `extract` on the slot pipeline (slot_pipeline.h): replaces the per-record loop of
// src/cmd_extract.rs:321-406 and :463-607 for FASTQ input. Mates of paired input are interleaved in
// the batch (record 2k / 2k+1 = pair k); a pair is never split over two batches.
#pragma once
#include "fastq_stream.h"
#include "slot_pipeline.h"

namespace mkh {

class FastqPipeline : public SlotPipeline {
public:
    // The readers are created by the caller (before the engines, so that reading and indexing the input
    // overlaps CUDA start-up); reader2 == nullptr: single-end.
    FastqPipeline(EngineSet& engines, std::unique_ptr<FastqChunkReader> reader1, std::unique_ptr<FastqChunkReader> reader2, mk_mode mode,
                  BatchConsumer consumer);
    ~FastqPipeline() override;
    static std::unique_ptr<FastqChunkReader> open_reader(const std::string& path, int n_files = 1);

protected:
    void begin() override;
    bool fill(PackedBatch& b) override;

private:
    bool paired_;
    std::unique_ptr<FastqChunkReader> rd_[2];
    std::shared_ptr<Chunk> cur_[2];
    size_t idx_[2] = {0, 0};
    std::string read_error_[2];  // what the reader of file f threw (decompression / read error); reported by fill()
    std::vector<CopyPool::Copy> copies_;  // the sequence copies of the batch being filled (carried out by copy_pool_)
};

}  // namespace mkh
