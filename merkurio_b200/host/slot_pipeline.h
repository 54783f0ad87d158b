// Input -> GPU pipeline shared by `extract` (FASTQ) and `tag` (SAM/BAM). Three stages, each on its
// own thread(s):
//
//   reader  (one thread per input file): read / inflate, index the records in place
//   packer  (one thread): copy the sequence bytes of successive records straight into a pinned slot
//           of an engine, until the slot is full            -> fill(), implemented per input format
//   driver  (the calling thread): submit the slot, wait for the oldest batch in flight, hand
//           (batch, result) to the command's consumer, return the slot to the packer
//
// A batch keeps references to the chunks its records live in, so that the consumer can name and
// write the records that need output without any per-record bookkeeping for those that do not.
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

namespace mkh {

// `count` consecutive records of one chunk; batch records [rec0, rec0 + count) of that input file.
struct BatchSeg {
    std::shared_ptr<const void> chunk;
    uint32_t first, count;
    uint32_t rec0;  // index among the batch's records of the same file
};

struct PackedBatch {
    int engine = 0;
    uint32_t slot = 0;
    uint8_t* seq = nullptr;
    uint64_t* off = nullptr;
    uint32_t* lens = nullptr;  // BAM4 batches only
    uint32_t n_records = 0;    // paired FASTQ: 2 x pairs, record 2k = mate 1 of pair k, record 2k+1 = its mate 2
    uint64_t n_units = 0;
    uint64_t n_bytes = 0;      // bytes used in seq
    uint64_t total_bases = 0;  // sum of the record lengths
    std::vector<BatchSeg> seg[2];  // per input file
    std::shared_ptr<void> extra;   // format-specific description of the batch's records (FASTA: its pieces)
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
    // (the chunk's reference count is only touched when a new run of records starts)
    template <class T>
    void add_to_seg(int f, const std::shared_ptr<T>& chunk, uint32_t idx, uint32_t rec_index) {
        std::vector<BatchSeg>& v = seg[f];
        if (!v.empty() && v.back().chunk.get() == chunk.get() && v.back().first + v.back().count == idx) v.back().count += 1;
        else v.push_back(BatchSeg{std::shared_ptr<const void>(chunk), idx, 1, rec_index});
    }
};

using BatchConsumer = std::function<void(const PackedBatch&, const mk_result&)>;

// Helper threads of the packer. Deciding which records go into a batch, and where, is one sequential pass over
// the record index (it has to be: a batch closes when the slot is full); copying the sequence bytes — most of the
// packer's time — is not. fill() notes the copies in a list and publishes the list's length every few thousand
// records; MERKURIO_PACK_THREADS - 1 workers (default: 4 threads or a quarter of the cores, the packer thread included)
// carry out blocks of the published part while the packer goes on deciding, and the packer joins them at the end.
class CopyPool {
public:
    struct Copy {
        const char* src;
        uint64_t dst;  // offset in the slot's sequence buffer
        uint32_t len;
    };
    explicit CopyPool(int threads);
    ~CopyPool();
    CopyPool(const CopyPool&) = delete;
    // A batch: begin(), any number of publish() calls with growing counts, finish(). `list` must not move in between
    // (the caller reserves it), and the sources must stay readable until finish() returns.
    void begin(uint8_t* base, const Copy* list);
    void publish(size_t n);  // list[0, n) is final
    void finish(size_t n);   // ... and n is all there is; returns when every copy is done
    int threads() const { return (int)workers_.size() + 1; }

private:
    static constexpr size_t kBlock = 2048;  // copies a thread takes at a time
    void work();
    bool take(std::unique_lock<std::mutex>& lk, size_t* lo, size_t* hi);  // a block of the published part, if any is left
    static void copy_range(uint8_t* base, const Copy* c, size_t lo, size_t hi);
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
    uint8_t* base_ = nullptr;
    const Copy* list_ = nullptr;
    size_t ready_ = 0, next_ = 0, done_ = 0;  // published / handed out / carried out
};

class SlotPipeline {
public:
    SlotPipeline(EngineSet& engines, mk_encoding enc, mk_mode mode, BatchConsumer consumer);
    virtual ~SlotPipeline();
    // Runs the whole input. Throws the input's error (parse error, unequal files) after everything before
    // it has been delivered, like the reference, which fails at the record it cannot read.
    void run();

protected:
    // Both run on the packer thread. fill() packs as many whole records as fit into b and returns false
    // when the input is exhausted and nothing was added; it sets input_done_ once nothing more will come.
    virtual void begin() {}
    virtual bool fill(PackedBatch& b) = 0;
    // Derived destructors call this first: the packer thread uses their members.
    void stop_packer();
    EngineSet& es_;
    mk_encoding enc_;
    mk_mode mode_;
    bool input_done_ = false;
    CopyPool copy_pool_;

private:
    void pack();
    BatchConsumer consumer_;
    std::thread packer_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::unique_ptr<PackedBatch>> free_, packed_;
    bool packer_done_ = false, stop_ = false;
    std::string packer_error_;
};

}  // namespace mkh
