#include "slot_pipeline.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace mkh {

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

static int pack_threads() {
    if (const char* s = std::getenv("MERKURIO_PACK_THREADS")) return std::max(1, std::atoi(s));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(4u, hw / 4));
}

CopyPool::CopyPool(int threads) {
    for (int t = 1; t < threads; ++t) workers_.emplace_back([this] { work(); });
}

CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
}

void CopyPool::copy_range(uint8_t* base, const Copy* c, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) std::memcpy(base + c[i].dst, c[i].src, c[i].len);
}

bool CopyPool::take(std::unique_lock<std::mutex>&, size_t* lo, size_t* hi) {
    if (next_ >= ready_) return false;
    *lo = next_;
    *hi = std::min(ready_, next_ + kBlock);
    next_ = *hi;
    return true;
}

void CopyPool::work() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        cv_.wait(lk, [&] { return stop_ || next_ < ready_; });
        if (stop_) return;
        size_t lo, hi;
        while (take(lk, &lo, &hi)) {
            uint8_t* base = base_;
            const Copy* list = list_;
            lk.unlock();
            copy_range(base, list, lo, hi);
            lk.lock();
            done_ += hi - lo;
            if (done_ == ready_) done_cv_.notify_all();
        }
    }
}

void CopyPool::begin(uint8_t* base, const Copy* list) {
    std::lock_guard<std::mutex> lk(mu_);
    base_ = base;
    list_ = list;
    ready_ = next_ = done_ = 0;
}

void CopyPool::publish(size_t n) {
    if (workers_.empty()) return;  // the packer does them all in finish()
    {
        std::lock_guard<std::mutex> lk(mu_);
        ready_ = n;
    }
    cv_.notify_all();
}

void CopyPool::finish(size_t n) {
    std::unique_lock<std::mutex> lk(mu_);
    ready_ = n;
    if (!workers_.empty() && next_ < ready_) {
        lk.unlock();
        cv_.notify_all();
        lk.lock();
    }
    size_t lo, hi;
    while (take(lk, &lo, &hi)) {  // the packer thread takes its share of what is left
        lk.unlock();
        copy_range(base_, list_, lo, hi);
        lk.lock();
        done_ += hi - lo;
    }
    done_cv_.wait(lk, [&] { return done_ == ready_; });
}

SlotPipeline::SlotPipeline(EngineSet& engines, mk_encoding enc, mk_mode mode, BatchConsumer consumer)
    : es_(engines), enc_(enc), mode_(mode), copy_pool_(pack_threads()), consumer_(std::move(consumer)) {
    // every slot of every engine, in the order the batches will use them
    const size_t G = es_.engines.size();
    for (uint32_t s = 0; s < es_.n_slots; ++s)
        for (size_t g = 0; g < G; ++g) {
            std::unique_ptr<PackedBatch> b(new PackedBatch);
            b->engine = (int)g;
            b->slot = s;
            if (mk_slot_buffers(es_.engines[g], s, &b->seq, &b->off, enc_ == MK_ENC_BAM4 ? &b->lens : nullptr) != 0)
                throw Error(std::string("GPU matching engine: ") + mk_last_error());
            free_.push_back(std::move(b));
        }
}

void SlotPipeline::stop_packer() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (packer_.joinable()) packer_.join();
}

SlotPipeline::~SlotPipeline() { stop_packer(); }

void SlotPipeline::pack() {
    try {
        begin();
        for (;;) {
            std::unique_ptr<PackedBatch> b;
            const double t0 = now_s();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !free_.empty() || stop_; });
                if (stop_) return;
                b = std::move(free_.front());
                free_.pop_front();
            }
            const double t1 = now_s();
            bool any = fill(*b);
            es_.t_pack_wait += t1 - t0;
            es_.t_pack += now_s() - t1;
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

void SlotPipeline::run() {
    const double t_run0 = now_s();
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
            if (inflight.empty()) {
                const double t0 = now_s();
                cv_.wait(lk, [this] { return !packed_.empty() || packer_done_; });
                es_.t_idle += now_s() - t0;
            }
            if (!packed_.empty()) {
                b = std::move(packed_.front());
                packed_.pop_front();
            } else if (packer_done_) {
                done = true;
            }
        }
        if (b) {
            if (b->n_records > 0) {
                const double t0 = now_s();
                const int rc_submit = mk_scan_submit(es_.engines[(size_t)b->engine], b->slot, b->n_records, b->n_units, enc_ == MK_ENC_BAM4 ? 1 : 0, enc_, mode_);
                es_.t_submit += now_s() - t0;
                if (rc_submit != 0) throw Error(std::string("GPU matching engine: ") + mk_last_error());
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
    es_.t_run += now_s() - t_run0;
    if (!packer_error_.empty()) throw Error(packer_error_);
    if (!error_chain.empty()) {
        Error e(error_chain.back());
        for (size_t i = error_chain.size() - 1; i-- > 0;) e = e.with_context(error_chain[i]);
        throw e;
    }
}

}  // namespace mkh
