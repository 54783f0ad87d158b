#include "fastq_pipeline.h"

#include <chrono>
#include <cstdlib>
#include <cstring>

namespace mkh {

namespace {
const char* kParseError = "Error during FASTQ/A record parsing.";
const char* kSecondCtx = "Error during FASTQ record parsing of second file. Do the two input files contain the same number of records?";
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

std::unique_ptr<FastqChunkReader> FastqPipeline::open_reader(const std::string& path) {
    const size_t chunk_bytes = std::getenv("MERKURIO_CHUNK_BYTES") ? (size_t)std::strtoull(std::getenv("MERKURIO_CHUNK_BYTES"), nullptr, 10)
                                                                   : (size_t)8 << 20;
    // up to 16 chunks are read ahead while the engines start up
    return std::unique_ptr<FastqChunkReader>(new FastqChunkReader(path, chunk_bytes, 16));
}

FastqPipeline::FastqPipeline(EngineSet& engines, std::unique_ptr<FastqChunkReader> reader1, std::unique_ptr<FastqChunkReader> reader2,
                             mk_mode mode, BatchConsumer consumer)
    : es_(engines), paired_((bool)reader2), mode_(mode), consumer_(std::move(consumer)) {
    // every slot of every engine, in the order the batches will use them
    const size_t G = es_.engines.size();
    for (uint32_t s = 0; s < es_.n_slots; ++s)
        for (size_t g = 0; g < G; ++g) {
            std::unique_ptr<PackedBatch> b(new PackedBatch);
            b->engine = (int)g;
            b->slot = s;
            if (mk_slot_buffers(es_.engines[g], s, &b->seq, &b->off, nullptr) != 0)
                throw Error(std::string("GPU matching engine: ") + mk_last_error());
            free_.push_back(std::move(b));
        }
    rd_[0] = std::move(reader1);
    rd_[1] = std::move(reader2);
}

FastqPipeline::~FastqPipeline() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (packer_.joinable()) packer_.join();
}

// Fill one slot with as many whole records (pairs) as fit. Runs on the packer thread.
bool FastqPipeline::fill(PackedBatch& b) {
    b.n_records = 0;
    b.n_units = 0;
    b.seg[0].clear();
    b.seg[1].clear();
    b.error_chain.clear();
    if (input_done_) return false;
    const int F = paired_ ? 2 : 1;
    const uint64_t cap = es_.max_bytes;
    const uint32_t max_rec = es_.max_records;
    uint32_t nrec[2] = {0, 0};
    auto fail = [&](std::vector<std::string> chain) {
        b.error_chain = std::move(chain);
        input_done_ = true;
    };
    // step file f to its next record; false at its end (cur_[f] == nullptr) or on a malformed record
    auto advance = [&](int f, bool* malformed) {
        *malformed = false;
        while (cur_[f] && idx_[f] == cur_[f]->recs.size()) {
            if (cur_[f]->failed) { *malformed = true; return false; }
            cur_[f] = rd_[f]->next();
            idx_[f] = 0;
        }
        return (bool)cur_[f];
    };
    auto add = [&](int f) {
        const Chunk& c = *cur_[f];
        const RecSpan& r = c.recs[idx_[f]];
        b.off[b.n_records] = b.n_units;
        if (r.seq_len) std::memcpy(b.seq + b.n_units, c.seq(r), r.seq_len);
        b.n_units += r.seq_len;
        b.n_records += 1;
        std::vector<BatchSeg>& v = b.seg[f];
        if (!v.empty() && v.back().chunk.get() == &c && v.back().first + v.back().count == idx_[f]) v.back().count += 1;
        else v.push_back(BatchSeg{cur_[f], (uint32_t)idx_[f], 1, nrec[f]});
        nrec[f] += 1;
        idx_[f] += 1;
    };
    for (;;) {
        bool bad = false;
        if (!advance(0, &bad)) {
            if (bad) {
                fail(paired_ ? std::vector<std::string>{"Error during FASTQ record parsing of first file.", kParseError}
                             : std::vector<std::string>{kParseError, kParseError});
            } else {
                if (paired_) {  // file 1 is exhausted: file 2 must be, too
                    bool bad2 = false;
                    if (advance(1, &bad2)) fail({"The two input files have a different number of records. Please provide valid paired-end read files."});
                    else if (bad2) fail({kParseError});
                }
                input_done_ = true;
            }
            break;
        }
        uint64_t need = cur_[0]->recs[idx_[0]].seq_len;
        if (paired_) {
            bool bad2 = false;
            if (!advance(1, &bad2)) {
                fail(bad2 ? std::vector<std::string>{kSecondCtx, kParseError} : std::vector<std::string>{kSecondCtx});
                break;
            }
            need += cur_[1]->recs[idx_[1]].seq_len;
        }
        if (need > cap)
            throw Error("a record of " + std::to_string(need) + " bases does not fit a batch (raise MERKURIO_BATCH_MB)");
        if (b.n_records + (uint32_t)F > max_rec || b.n_units + need > cap) break;  // full: the record opens the next batch
        add(0);
        if (paired_) add(1);
    }
    b.off[b.n_records] = b.n_units;
    return b.n_records > 0 || !b.error_chain.empty();
}

void FastqPipeline::pack() {
    try {
        cur_[0] = rd_[0]->next();
        if (paired_) cur_[1] = rd_[1]->next();
        for (;;) {
            std::unique_ptr<PackedBatch> b;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !free_.empty() || stop_; });
                if (stop_) return;
                b = std::move(free_.front());
                free_.pop_front();
            }
            bool any = fill(*b);
            std::lock_guard<std::mutex> lk(mu_);
            if (any) packed_.push_back(std::move(b));
            else free_.push_front(std::move(b));
            if (!any || input_done_) { packer_done_ = true; cv_.notify_all(); return; }
            cv_.notify_all();
        }
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(mu_);
        packer_error_ = e.what();
        packer_done_ = true;
        cv_.notify_all();
    }
}

void FastqPipeline::run() {
    packer_ = std::thread([this] { pack(); });
    std::deque<std::unique_ptr<PackedBatch>> inflight;
    std::vector<std::string> error_chain;
    auto consume_oldest = [&] {
        std::unique_ptr<PackedBatch> b = std::move(inflight.front());
        inflight.pop_front();
        mk_result res{};
        es_.wait(b->engine, b->slot, &res);
        const double t0 = now_s();
        consumer_(*b, res);
        if (!b->error_chain.empty()) error_chain = b->error_chain;
        b->seg[0].clear();  // drop the chunk references before the slot goes back
        b->seg[1].clear();
        es_.t_deliver += now_s() - t0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            free_.push_back(std::move(b));
        }
        cv_.notify_all();
    };
    for (;;) {
        std::unique_ptr<PackedBatch> b;
        bool done = false;
        {
            std::unique_lock<std::mutex> lk(mu_);
            // with batches in flight do not block on the packer: consuming them is what frees its slots
            if (inflight.empty()) cv_.wait(lk, [this] { return !packed_.empty() || packer_done_; });
            if (!packed_.empty()) {
                b = std::move(packed_.front());
                packed_.pop_front();
            } else if (packer_done_) {
                done = true;
            }
        }
        if (b) {
            if (b->n_records > 0) {
                if (mk_scan_submit(es_.engines[(size_t)b->engine], b->slot, b->n_records, b->n_units, 0, MK_ENC_ASCII, mode_) != 0)
                    throw Error(std::string("GPU matching engine: ") + mk_last_error());
                inflight.push_back(std::move(b));
            } else {
                // nothing but the input's error
                while (!inflight.empty()) consume_oldest();
                error_chain = b->error_chain;
            }
            continue;
        }
        if (!inflight.empty()) { consume_oldest(); continue; }
        if (done) break;
    }
    if (packer_.joinable()) packer_.join();
    if (!packer_error_.empty()) throw Error(packer_error_);
    if (!error_chain.empty()) {
        Error e(error_chain.back());
        for (size_t i = error_chain.size() - 1; i-- > 0;) e = e.with_context(error_chain[i]);
        throw e;
    }
}

}  // namespace mkh
