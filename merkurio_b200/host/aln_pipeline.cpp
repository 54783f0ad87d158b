#include "aln_pipeline.h"

#include <cstring>

namespace mkh {

AlnPipeline::AlnPipeline(EngineSet& engines, std::unique_ptr<AlnChunkReader> reader, mk_mode mode, BatchConsumer consumer)
    : SlotPipeline(engines, MK_ENC_BAM4, mode, std::move(consumer)), rd_(std::move(reader)) {
    for (int c = 0; c < 256; ++c) code_lut_[c] = nibble_of_sam_char((char)c);
    pair_lut_.resize(65536);
    for (int hi = 0; hi < 256; ++hi)
        for (int lo = 0; lo < 256; ++lo) pair_lut_[(size_t)(hi << 8 | lo)] = (uint8_t)((code_lut_[lo] << 4) | code_lut_[hi]);
}

AlnPipeline::~AlnPipeline() { stop_packer(); }

void AlnPipeline::begin() {
    try {
        cur_ = rd_->next();
    } catch (const Error& e) {  // raised by the first fill(), with the record path's context
        begin_error_ = e.what();
    }
}

bool AlnPipeline::fill(PackedBatch& b) {
    b.n_records = 0;
    b.n_units = b.n_bytes = b.total_bases = 0;
    b.seg[0].clear();
    b.seg[1].clear();
    b.error_chain.clear();
    if (input_done_) return false;
    if (!begin_error_.empty()) {
        b.error_chain = {std::string("Error during ") + (rd_->is_bam() ? "BAM" : "SAM") + " record parsing: " + begin_error_};
        b.off[0] = 0;
        input_done_ = true;
        return true;
    }
    const uint64_t cap = es_.max_bytes;
    const uint32_t max_rec = es_.max_records;
    for (;;) {
        while (cur_ && idx_ == cur_->recs.size()) {
            if (!cur_->error.empty()) {
                b.error_chain = {std::string("Error during ") + (cur_->bam ? "BAM" : "SAM") + " record parsing: " + cur_->error};
                input_done_ = true;
                break;
            }
            try {
                cur_ = rd_->next();
            } catch (const Error& e) {
                // read / decompression error: what has been packed so far still goes out, then the error is raised
                // with the record path's context
                cur_ = nullptr;
                b.error_chain = {std::string("Error during ") + (rd_->is_bam() ? "BAM" : "SAM") + " record parsing: " + e.what()};
                input_done_ = true;
                break;
            }
            idx_ = 0;
        }
        if (input_done_) break;
        if (!cur_) { input_done_ = true; break; }
        const AlnSpan& r = cur_->recs[idx_];
        const uint64_t nbytes = ((uint64_t)r.l_seq + 1) / 2;
        if (nbytes > cap) throw Error("record with " + std::to_string(r.l_seq) + " bases exceeds the batch size (set MERKURIO_BATCH_MB)");
        if (b.n_records + 1 > max_rec || b.n_bytes + nbytes > cap) break;  // full
        b.off[b.n_records] = b.n_units;  // even: records are byte aligned
        b.lens[b.n_records] = r.l_seq;
        uint8_t* dst = b.seq + b.n_bytes;
        const char* src = cur_->data.data() + r.seq_off;
        if (cur_->bam) {
            if (nbytes) std::memcpy(dst, src, nbytes);
        } else {
            const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
            uint32_t i = 0;
            const uint8_t* lut = pair_lut_.data();  // one look-up per packed byte
            for (; i + 1 < r.l_seq; i += 2) *dst++ = lut[(size_t)s[i] | (size_t)s[i + 1] << 8];
            if (i < r.l_seq) *dst = (uint8_t)(code_lut_[s[i]] << 4);
        }
        b.n_bytes += nbytes;
        b.n_units += nbytes * 2;
        b.total_bases += r.l_seq;
        b.add_to_seg(0, cur_, (uint32_t)idx_, b.n_records);
        b.n_records += 1;
        idx_ += 1;
    }
    b.off[b.n_records] = b.n_units;
    return b.n_records > 0 || !b.error_chain.empty();
}

}  // namespace mkh
