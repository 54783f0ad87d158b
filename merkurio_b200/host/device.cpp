#include "device.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace mkh {

static void check(int rc) {
    if (rc != 0) throw Error(std::string("GPU matching engine: ") + mk_last_error());
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// MERKURIO_TIMING: when this process image was loaded, and a line at the very end of exit() (registered
// before the CUDA runtime registers its own handler, so it runs after the context has been torn down).
static const double g_loaded_at = now_s();
bool g_leave_engines_to_process_exit = false;

static void report_process_time() { std::fprintf(stderr, "[merkurio] process: %.3f s from load to the end of exit()\n", now_s() - g_loaded_at); }

void report_process_time_if_asked() {
    if (std::getenv("MERKURIO_TIMING")) report_process_time();
}

static uint64_t env_u64(const char* name, uint64_t dflt) {
    const char* s = std::getenv(name);
    if (!s || !*s) return dflt;
    return std::strtoull(s, nullptr, 10);
}

EngineSet::EngineSet(const std::vector<std::string>& patterns, bool case_insensitive, uint32_t default_batch_mb) {
    t_start = now_s();
    if (std::getenv("MERKURIO_TIMING")) {
        static bool once = false;
        if (!once) { once = true; std::atexit(report_process_time); }
        std::fprintf(stderr, "[merkurio] process: engines requested %.3f s after load\n", t_start - g_loaded_at);
    }
    std::string blob;
    std::vector<uint32_t> off{0};
    for (auto& p : patterns) {
        blob += p;
        off.push_back((uint32_t)blob.size());
        max_pattern_len = std::max<uint32_t>(max_pattern_len, (uint32_t)p.size());
    }
    mk_patterns mp{reinterpret_cast<const uint8_t*>(blob.data()), off.data(), (uint32_t)patterns.size()};
    // MERKURIO_GPUS=N: one engine on each of the devices 0..N-1. MERKURIO_DEVICES=a,b,...: one engine per entry on
    // exactly those devices (an ordinal may repeat: "0,0" drives two engines on one GPU, which is how the
    // round-robin deal and the in-order merge are tested on a single-GPU box).
    std::vector<int> devices;
    if (const char* dl = std::getenv("MERKURIO_DEVICES")) {
        for (const char* c = dl; *c;) {
            char* end = nullptr;
            long v = std::strtol(c, &end, 10);
            if (end == c || v < 0) throw Error(std::string("MERKURIO_DEVICES: cannot read the device list '") + dl + "'");
            devices.push_back((int)v);
            c = (*end == ',') ? end + 1 : end;
            if (*end && *end != ',') throw Error(std::string("MERKURIO_DEVICES: cannot read the device list '") + dl + "'");
        }
    }
    if (devices.empty()) {
        const int n_gpus = std::max((int)env_u64("MERKURIO_GPUS", 1), 1);
        for (int g = 0; g < n_gpus; ++g) devices.push_back(g);
    }
    max_bytes = env_u64("MERKURIO_BATCH_MB", default_batch_mb) << 20;
    if (uint64_t b = env_u64("MERKURIO_BATCH_BYTES", 0)) max_bytes = b;  // tests: tiny batches, many pieces
    if (max_bytes < (uint64_t)4 * max_pattern_len + 16384) max_bytes = (uint64_t)4 * max_pattern_len + 16384;
    // records per batch: sized for reads of >= 32 bases (shorter ones just close their batch early)
    max_records = (uint32_t)std::min<uint64_t>(max_bytes / 32 + 1024, 1u << 26);
    n_slots = (uint32_t)env_u64("MERKURIO_SLOTS", 3);
    if (n_slots < 1) n_slots = 1;
    // The seed tables are built once (mk_tables_create starts that on host threads) and shared by all engines,
    // which are created side by side: a CUDA context takes a second or more to create, one after the other that
    // is most of an 8-GPU run.
    mk_tables* tables = nullptr;
    check(mk_tables_create(&mp, case_insensitive ? 1 : 0, &tables));
    const int n_engines = (int)devices.size();
    engines.assign((size_t)n_engines, nullptr);
    std::vector<std::string> failed((size_t)n_engines);
    auto create = [&](int g) {
        mk_config cfg{};
        cfg.device = devices[(size_t)g];
        cfg.case_insensitive = case_insensitive ? 1 : 0;
        cfg.n_slots = n_slots;
        cfg.max_batch_records = max_records;
        cfg.max_batch_bytes = max_bytes;
        cfg.hit_capacity = 0;
        mk_engine* e = nullptr;
        if (mk_engine_create_shared(tables, &cfg, &e) != 0) failed[(size_t)g] = std::string("GPU matching engine: ") + mk_last_error();  // (thread-local text)
        engines[(size_t)g] = e;
    };
    if (n_engines == 1) {
        create(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < n_engines; ++g) th.emplace_back(create, g);
        for (auto& t : th) t.join();
    }
    mk_tables_destroy(tables);  // the engines hold their own references
    for (int g = 0; g < n_engines; ++g)
        if (!failed[(size_t)g].empty()) {
            for (mk_engine* e : engines) mk_engine_destroy(e);
            engines.clear();
            throw Error(failed[(size_t)g]);
        }
    t_setup = now_s() - t_start;
}

EngineSet::~EngineSet() {
    const double t_d0 = now_s();
    if (!g_leave_engines_to_process_exit)
        for (mk_engine* e : engines) mk_engine_destroy(e);
    const double t_destroy = now_s() - t_d0;
    if (std::getenv("MERKURIO_TIMING"))
        std::fprintf(stderr, "[merkurio] engine setup %.3f s, %llu batches, %llu records, %.3f Gbases, waited %.3f s for the GPU, "
                     "device time %.3f s, delivering results %.3f s, packer %.3f s busy (incl. waiting for input) + %.3f s waiting for a slot, pipeline %.3f s (submitting %.3f s, idle %.3f s), engine teardown %.3f s, total %.3f s\n", t_setup, (unsigned long long)n_batches,
                     (unsigned long long)n_records, (double)n_bases / 1e9, t_wait, (double)device_ns / 1e9, t_deliver, t_pack, t_pack_wait, t_run, t_submit, t_idle, t_destroy, now_s() - t_start);
}

void EngineSet::wait(int engine, uint32_t slot, mk_result* out) {
    double t0 = now_s();
    check(mk_scan_wait(engines[(size_t)engine], slot, out));
    t_wait += now_s() - t0;
    device_ns += out->device_ns;
    n_records += out->n_records;
    n_bases += out->bases_scanned;
    n_batches += 1;
}

Scanner::Scanner(EngineSet& engines, mk_encoding enc, mk_mode mode, RecordCallback cb)
    : es_(engines), engines_(engines.engines), enc_(enc), mode_(mode), cb_(std::move(cb)) {
    n_slots_ = es_.n_slots;
    max_records_ = es_.max_records;
    max_pattern_len_ = es_.max_pattern_len;
    max_bytes_ = es_.max_bytes;
}

Scanner::~Scanner() {}

void Scanner::open_batch() {
    // batch i goes to engine i % G, slot (i / G) % S; at most G*S batches are in flight
    const uint64_t G = engines_.size();
    while (inflight_.size() >= G * n_slots_) consume_oldest();
    // batches are recycled so that their vectors keep their (already touched) storage
    if (!spare_.empty()) {
        open_ = std::move(spare_.back());
        spare_.pop_back();
        open_->n_records = 0;
        open_->n_units = open_->n_bytes = 0;
        open_->pieces.clear();
        open_->metas.clear();
    } else {
        open_.reset(new Batch);
    }
    open_->engine = (int)(batch_seq_ % G);
    open_->slot = (uint32_t)((batch_seq_ / G) % n_slots_);
    ++batch_seq_;
    check(mk_slot_buffers(engines_[(size_t)open_->engine], open_->slot, &open_->seq, &open_->off,
                          enc_ == MK_ENC_BAM4 ? &open_->lens : nullptr));
}

void Scanner::submit_open() {
    if (!open_) return;
    Batch& b = *open_;
    if (b.n_records == 0) { --batch_seq_; open_.reset(); return; }
    b.off[b.n_records] = b.n_units;
    check(mk_scan_submit(engines_[(size_t)b.engine], b.slot, b.n_records, b.n_units, enc_ == MK_ENC_BAM4 ? 1 : 0, enc_, mode_));
    inflight_.push_back(std::move(open_));
}

void Scanner::add_record(const char* seq, size_t len, RecMeta&& meta) {
    meta.len = (uint32_t)len;
    const uint64_t overlap = max_pattern_len_ ? max_pattern_len_ - 1 : 0;
    uint64_t pos = 0;  // first base of the record not yet covered by a piece's own range
    bool first = true;
    for (;;) {
        if (!open_) open_batch();
        Batch* b = open_.get();
        uint64_t lead = first ? 0 : std::min(overlap, pos);
        uint64_t room = max_bytes_ - b->n_bytes;
        // a piece carries its overlap plus new bases; close the batch if the rest of the record does not
        // fit and there is not even room for a useful piece
        if (b->n_records >= max_records_ || (b->n_records > 0 && room < std::min<uint64_t>(len - pos + lead, lead + 4096))) {
            submit_open();
            continue;
        }
        uint64_t take = std::min<uint64_t>(len - pos, room - lead);
        Piece pc;
        pc.first = first;
        pc.base = pos - lead;
        pc.own_from = (uint32_t)lead;
        pc.last = (pos + take == len);
        b->off[b->n_records] = b->n_units;
        if (lead + take) std::memcpy(b->seq + b->n_bytes, seq + pos - lead, lead + take);
        b->n_bytes += lead + take;
        b->n_units += lead + take;
        b->n_records++;
        b->pieces.push_back(pc);
        if (first) b->metas.push_back(std::move(meta));
        pos += take;
        first = false;
        if (pc.last) break;
        submit_open();
    }
}

void Scanner::add_record_packed(const uint8_t* packed, uint32_t l_seq, RecMeta&& meta) {
    meta.len = l_seq;
    uint64_t nbytes = ((uint64_t)l_seq + 1) / 2;
    if (nbytes > max_bytes_) throw Error("record with " + std::to_string(l_seq) + " bases exceeds the batch size (set MERKURIO_BATCH_MB)");
    if (!open_) open_batch();
    if (open_->n_records >= max_records_ || open_->n_bytes + nbytes > max_bytes_) {
        submit_open();
        open_batch();
    }
    Batch* b = open_.get();
    b->off[b->n_records] = b->n_units;  // even: records are byte aligned
    b->lens[b->n_records] = l_seq;
    if (nbytes) std::memcpy(b->seq + b->n_bytes, packed, nbytes);
    b->n_bytes += nbytes;
    b->n_units += nbytes * 2;
    b->n_records++;
    b->pieces.push_back(Piece{true, true, 0, 0});
    b->metas.push_back(std::move(meta));
}

void Scanner::deliver_piece(Batch& b, uint32_t r, bool flag, const mk_hit* hits, size_t n) {
    const Piece& pc = b.pieces[r];
    if (pc.first) {
        cur_hits_.clear();
        cur_found_ = false;
    }
    for (size_t i = 0; i < n; ++i) {
        const mk_hit& h = hits[i];
        if (mode_ == MK_MODE_ALL_HITS) {
            if (!pc.first && (uint64_t)h.start + h.len <= pc.own_from) continue;  // owned by the previous piece
            cur_hits_.push_back(RecHit{pc.base + h.start, h.pattern, h.len});
        } else {
            cur_hits_.push_back(RecHit{0, h.pattern, h.len});
        }
    }
    cur_found_ = cur_found_ || flag;
}

void Scanner::consume_oldest() {
    std::unique_ptr<Batch> bp = std::move(inflight_.front());
    inflight_.pop_front();
    Batch& b = *bp;
    mk_result res{};
    es_.wait(b.engine, b.slot, &res);
    const double t_in = now_s();
    size_t hi = 0, meta_i = 0;
    for (uint32_t r = 0; r < b.n_records; ++r) {
        size_t h0 = hi;
        while (hi < res.n_hits && res.hits[hi].record == r) ++hi;
        bool flag = (res.record_flags[r >> 6] >> (r & 63)) & 1;
        const Piece& pc = b.pieces[r];
        if (pc.first) cur_meta_ = std::move(b.metas[meta_i++]);
        deliver_piece(b, r, flag, res.hits ? res.hits + h0 : nullptr, hi - h0);
        if (pc.last) {
            if (mode_ == MK_MODE_PATTERN_SET && cur_hits_.size() > 1) {  // pieces of one record may repeat a pattern
                std::sort(cur_hits_.begin(), cur_hits_.end(), [](const RecHit& x, const RecHit& y) { return x.pattern < y.pattern; });
                cur_hits_.erase(std::unique(cur_hits_.begin(), cur_hits_.end(),
                                            [](const RecHit& x, const RecHit& y) { return x.pattern == y.pattern; }),
                                cur_hits_.end());
            }
            cb_(cur_meta_, cur_found_, cur_hits_);
        }
    }
    b.metas.clear();
    spare_.push_back(std::move(bp));
    es_.t_deliver += now_s() - t_in;
}

void Scanner::finish() {
    submit_open();
    while (!inflight_.empty()) consume_oldest();
}

}  // namespace mkh
