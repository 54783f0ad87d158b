#include "fastq_pipeline.h"

#include <cstdlib>
#include <cstring>

namespace mkh {

namespace {
const char* kParseError = "Error during FASTQ/A record parsing.";
const char* kSecondCtx = "Error during FASTQ record parsing of second file. Do the two input files contain the same number of records?";
}  // namespace

std::unique_ptr<FastqChunkReader> FastqPipeline::open_reader(const std::string& path, int n_files) {
    const size_t chunk_bytes = std::getenv("MERKURIO_CHUNK_BYTES") ? (size_t)std::strtoull(std::getenv("MERKURIO_CHUNK_BYTES"), nullptr, 10)
                                                                   : (size_t)8 << 20;
    // read ahead (by byte budget) while the engines start up
    return std::unique_ptr<FastqChunkReader>(new FastqChunkReader(path, chunk_bytes, prefetch_depth(chunk_bytes, n_files)));
}

FastqPipeline::FastqPipeline(EngineSet& engines, std::unique_ptr<FastqChunkReader> reader1, std::unique_ptr<FastqChunkReader> reader2,
                             mk_mode mode, BatchConsumer consumer)
    : SlotPipeline(engines, MK_ENC_ASCII, mode, std::move(consumer)), paired_((bool)reader2) {
    rd_[0] = std::move(reader1);
    rd_[1] = std::move(reader2);
}

FastqPipeline::~FastqPipeline() { stop_packer(); }

void FastqPipeline::begin() {
    for (int f = 0; f < (paired_ ? 2 : 1); ++f) {
        try {
            cur_[f] = rd_[f]->next();
        } catch (const Error& e) {  // surfaces in fill(), where the error gets the record path's context
            read_error_[f] = e.what();
        }
    }
}

// Fill one slot with as many whole records (pairs) as fit. Runs on the packer thread.
bool FastqPipeline::fill(PackedBatch& b) {
    b.n_records = 0;
    b.n_units = b.n_bytes = b.total_bases = 0;
    b.seg[0].clear();
    b.seg[1].clear();
    b.error_chain.clear();
    if (input_done_) return false;
    const int F = paired_ ? 2 : 1;
    const uint64_t cap = es_.max_bytes;
    const uint32_t max_rec = es_.max_records;
    uint32_t nrec[2] = {0, 0};
    // the copies are carried out by the pool while this thread goes on choosing records; the list must not move
    copies_.clear();
    copies_.reserve((size_t)max_rec + 2);
    copy_pool_.begin(b.seq, copies_.data());
    size_t published = 0;
    auto fail = [&](std::vector<std::string> chain) {
        b.error_chain = std::move(chain);
        input_done_ = true;
    };
    // step file f to its next record; false at its end (cur_[f] == nullptr), on a malformed record or when the
    // input cannot be read any further (*why: what the record-by-record reader would have thrown there)
    auto advance = [&](int f, bool* malformed, std::string* why) {
        *malformed = false;
        if (!read_error_[f].empty()) { *malformed = true; *why = read_error_[f]; return false; }
        while (cur_[f] && idx_[f] == cur_[f]->recs.size()) {
            if (cur_[f]->failed) { *malformed = true; *why = kParseError; return false; }
            try {
                cur_[f] = rd_[f]->next();
            } catch (const Error& e) {  // decompression / read error: reported like a parse error, with its own text
                cur_[f] = nullptr;
                read_error_[f] = e.what();
                *malformed = true;
                *why = read_error_[f];
                return false;
            }
            idx_[f] = 0;
        }
        return (bool)cur_[f];
    };
    auto add = [&](int f) {
        const Chunk& c = *cur_[f];
        const RecSpan& r = c.recs[idx_[f]];
        b.off[b.n_records] = b.n_units;
        if (r.seq_len) copies_.push_back(CopyPool::Copy{c.seq(r), b.n_units, r.seq_len});
        b.n_units += r.seq_len;
        b.n_records += 1;
        b.add_to_seg(f, cur_[f], (uint32_t)idx_[f], nrec[f]);
        nrec[f] += 1;
        idx_[f] += 1;
    };
    for (;;) {
        bool bad = false;
        std::string why;
        if (!advance(0, &bad, &why)) {
            if (bad) {
                fail(paired_ ? std::vector<std::string>{"Error during FASTQ record parsing of first file.", why}
                             : std::vector<std::string>{kParseError, why});
            } else {
                if (paired_) {  // file 1 is exhausted: file 2 must be, too
                    bool bad2 = false;
                    if (advance(1, &bad2, &why)) fail({"The two input files have a different number of records. Please provide valid paired-end read files."});
                    else if (bad2) fail({why});
                }
                input_done_ = true;
            }
            break;
        }
        uint64_t need = cur_[0]->recs[idx_[0]].seq_len;
        if (paired_) {
            bool bad2 = false;
            if (!advance(1, &bad2, &why)) {
                fail(bad2 ? std::vector<std::string>{kSecondCtx, why} : std::vector<std::string>{kSecondCtx});
                break;
            }
            need += cur_[1]->recs[idx_[1]].seq_len;
        }
        if (need > cap) {
            copy_pool_.finish(copies_.size());  // no copy is left running behind the exception
            throw Error("a record of " + std::to_string(need) + " bases does not fit a batch (raise MERKURIO_BATCH_MB)");
        }
        if (b.n_records + (uint32_t)F > max_rec || b.n_units + need > cap) break;  // full: the record opens the next batch
        add(0);
        if (paired_) add(1);
        if (copies_.size() >= published + 4096) {
            published = copies_.size();
            copy_pool_.publish(published);
        }
    }
    b.off[b.n_records] = b.n_units;
    b.n_bytes = b.total_bases = b.n_units;
    copy_pool_.finish(copies_.size());  // the chunks the copies read from are held by b.seg
    copies_.clear();
    return b.n_records > 0 || !b.error_chain.empty();
}

}  // namespace mkh
