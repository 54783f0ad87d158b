// SAM / BAM input of `tag` on the slot pipeline (slot_pipeline.h): replaces the record loops of
// src/cmd_tag.rs:530-557,585-612. BAM sequences are copied into the batch in their 4-bit packing
// (the device scans them directly); SAM text is packed to the same codes.
#pragma once
#include "aln_stream.h"
#include "slot_pipeline.h"

namespace mkh {

class AlnPipeline : public SlotPipeline {
public:
    AlnPipeline(EngineSet& engines, std::unique_ptr<AlnChunkReader> reader, mk_mode mode, BatchConsumer consumer);
    ~AlnPipeline() override;

protected:
    void begin() override;
    bool fill(PackedBatch& b) override;

private:
    std::unique_ptr<AlnChunkReader> rd_;
    std::shared_ptr<AlnChunk> cur_;
    size_t idx_ = 0;
    std::string begin_error_;             // what the reader threw before the first chunk
    uint8_t code_lut_[256];               // SAM character -> BAM code
    std::vector<uint8_t> pair_lut_;       // two SAM characters (first in the low byte) -> one packed byte
};

}  // namespace mkh
