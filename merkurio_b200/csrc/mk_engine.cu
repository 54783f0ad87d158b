// C ABI of the matching engine (include/merkurio_cuda.h): table upload, batch staging on CUDA
// streams, kernel dispatch, overflow handling. No CPU matching path exists in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <future>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "mk_scan.cuh"
#include "mk_sort.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(MK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    cudaError_t upload(const std::vector<T>& v) {
        cudaError_t e = ensure(v.size() ? v.size() : 1);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
};

template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) cap = n;
        return e;
    }
};

struct DeviceTables {
    bool built = false;
    const mk::Tables* host = nullptr;  // owned by the engine's mk_tables
    DevBuf<uint32_t> filter, filter2, postings, pat_off;
    DevBuf<mk::SeedSlot> slots;
    DevBuf<uint8_t> pat_bytes;
    uint64_t bytes() const {
        return host->slots.size() * sizeof(mk::SeedSlot) + host->postings.size() * 4 + host->pat_bytes.size() +
               host->pat_off.size() * 4 + host->filter.size() * 4 + host->filter2.size() * 4;
    }
};

// Everything one in-flight batch needs on the device and for its results on the host.
struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_scan = nullptr, ev_verify = nullptr, ev_end = nullptr;
    DevBuf<uint32_t> flags;                  // bitmap, 2 words per 64 records
    DevBuf<mk::RawHit> raw_a, raw_b;
    DevBuf<mk_hit> out;
    DevBuf<uint32_t> heads, radix_table, radix_totals;  // heads: per-tile head counts of the pair de-duplication
    DevBuf<uint32_t> buckets;                // bucket sort: counts [kBuckets], starts [kBuckets + 1], cursors [kBuckets]
    DevBuf<uint2> cand;                      // candidate list handed from the scan to the verify kernel
    uint64_t cand_cap = 0;
    DevBuf<unsigned long long> cta_clock;    // MK_CTA_CLOCKS=1: per-CTA start / end times of the scan kernel (diagnostics)
    DevBuf<unsigned long long> counters;     // [0] hits appended, [1] distinct pairs, [2] candidates
    PinBuf<unsigned long long> h_counters;
    PinBuf<uint64_t> h_flags;
    PinBuf<mk_hit> h_hits;
    uint64_t hit_cap = 0;
    // description of the batch in flight
    bool busy = false;
    bool prefer_radix = false;  // a bucket of the bucket sort overflowed once: this workspace sorts with the radix passes
    bool used_buckets = false;  // the batch in flight was sorted with the bucket sort
    uint32_t key_bits = 0;
    const void* d_seq = nullptr;
    const unsigned long long* d_off = nullptr;
    const uint32_t* d_lens = nullptr;
    uint32_t n_records = 0;
    uint64_t n_units = 0;
    mk_encoding enc = MK_ENC_ASCII;
    mk_mode mode = MK_MODE_FLAG;
    bool fetch = true;
    ~Workspace() {
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_scan) cudaEventDestroy(ev_scan);
        if (ev_verify) cudaEventDestroy(ev_verify);
        if (ev_end) cudaEventDestroy(ev_end);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct Slot {
    Workspace ws;
    PinBuf<uint8_t> h_seq;
    PinBuf<uint64_t> h_off;
    PinBuf<uint32_t> h_lens;
    DevBuf<uint8_t> d_seq;
    DevBuf<unsigned long long> d_off;
    DevBuf<uint32_t> d_lens;
};

}  // namespace

// The host side of a query set: pattern bytes and the seed tables of both encodings, built once (on demand, on
// host threads for large sets) and shared by every engine created from it — one engine per GPU uploads its own
// device copy. Reference-counted: the engines hold it alive after mk_tables_destroy.
struct mk_tables {
    mk::PatternSet ps;
    std::mutex mu;
    mk::Tables host[2];
    bool built[2] = {false, false};
    std::future<mk::Tables> pending[2];
    std::atomic<int> refs{1};

    void start_async() {  // large query sets take seconds to index: start now, whichever encoding will be scanned
        for (int enc = 0; enc < 2; ++enc)
            pending[enc] = std::async(std::launch::async, [this, enc] { return mk::build_tables(ps, enc); });
    }
    // the tables of one encoding (thread-safe; throws what the builder throws)
    const mk::Tables* get(int enc) {
        std::lock_guard<std::mutex> lock(mu);
        if (!built[enc]) {
            if (pending[enc].valid()) host[enc] = pending[enc].get();
            else host[enc] = mk::build_tables(ps, enc);
            built[enc] = true;
        }
        return &host[enc];
    }
    void release() {
        if (refs.fetch_sub(1) == 1) {
            for (auto& f : pending) if (f.valid()) f.wait();
            delete this;
        }
    }
};

struct mk_engine {
    int device = 0;
    int sm_count = 0;
    mk_config cfg{};
    mk_tables* tab = nullptr;
    DeviceTables tables[2];
    ~mk_engine() { if (tab) tab->release(); }
    DevBuf<uint32_t> tie_rank;
    std::vector<std::unique_ptr<Slot>> slots;
    Workspace direct;  // mk_scan_device
    cudaEvent_t last_device_submit = nullptr;  // end of the latest mk_scan_device_submit batch (its slot's ev_end)
};

namespace {

using ScanKernel = void (*)(const mk::ScanParams);

// Launch shape of a scan kernel: threads per CTA (one CTA per SM) and 16-byte vectors per tile.
struct ScanLaunch {
    ScanKernel fn;
    int threads;
    int tile_vecs;
};

// Stride-16 scan: U vectors per lane and tile (two tiles in flight), T threads per CTA.
// MK_TUNE_U / MK_TUNE_T override the defaults (used by scripts/tune_scan.py only).
#ifdef MK_TUNE_BUILD
int tune_env(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s ? std::atoi(s) : dflt;
}
#endif

#ifdef MK_TUNE_BUILD
// Every launch shape scripts/tune_scan.py sweeps (build with -DMK_TUNE_BUILD; ~100 extra kernels).
template <int ENC, int FMODE, int T>
ScanKernel pick_d16_u(int u, bool v8) {
    if (v8) {
        switch (u) {
            case 8: return mk::mk_scan_d16<ENC, FMODE, 8, T, true>;
            case 4: return mk::mk_scan_d16<ENC, FMODE, 4, T, true>;
            default: return mk::mk_scan_d16<ENC, FMODE, 2, T, true>;
        }
    }
    switch (u) {
        case 8: return mk::mk_scan_d16<ENC, FMODE, 8, T, false>;
        case 4: return mk::mk_scan_d16<ENC, FMODE, 4, T, false>;
        case 3: return mk::mk_scan_d16<ENC, FMODE, 3, T, false>;
        default: return mk::mk_scan_d16<ENC, FMODE, 2, T, false>;
    }
}
#endif

// MK_D16_SHAPE (tuning): launch shape of the stride-16 scan, see pick_by_d
int d16_shape() {
    const char* v = std::getenv("MK_D16_SHAPE");  // read per batch: a tuning process sweeps the shapes
    return v ? std::atoi(v) : 0;
}

template <int ENC, int FMODE>
ScanLaunch pick_by_d(uint32_t d) {
    if (d == 16) {
#ifdef MK_TUNE_BUILD
        int u = tune_env("MK_TUNE_U", 4), t = tune_env("MK_TUNE_T", 896);
        bool v8 = tune_env("MK_TUNE_V8", 0) != 0;
        if (u != 2 && u != 3 && u != 4 && u != 8) u = 2;
        if (v8 && u == 3) u = 4;
        switch (t) {
            case 512: return {pick_d16_u<ENC, FMODE, 512>(u, v8), 512, u * 32};
            case 768: return {pick_d16_u<ENC, FMODE, 768>(u, v8), 768, u * 32};
            case 896: return {pick_d16_u<ENC, FMODE, 896>(u, v8), 896, u * 32};
            default: return {pick_d16_u<ENC, FMODE, 1024>(u, v8), 1024, u * 32};
        }
#else
        // measured best shapes. ASCII: 4 vectors per lane and tile, 896 threads, 16-byte loads. BAM4 (bound by its shared-memory probes, not by
        // HBM): 768 threads, 2.8 % faster on BASELINE cfg4 (0.631 -> 0.607 ms). MK_D16_SHAPE=1 / 2 force 768 / 896.
        const int shape = d16_shape();
        if (shape == 1 || (shape == 0 && ENC == MK_ENC_BAM4)) return {mk::mk_scan_d16<ENC, FMODE, 4, 768, false>, 768, 4 * 32};
        return {mk::mk_scan_d16<ENC, FMODE, 4, 896, false>, 896, 4 * 32};
#endif
    }
    switch (d) {
        case 8: return {mk::mk_scan_ord<ENC, 8, FMODE, 4>, mk::kScanThreads, 4 * 32};
        case 4: return {mk::mk_scan_ord<ENC, 4, FMODE, 4>, mk::kScanThreads, 4 * 32};
        case 2: return {mk::mk_scan_ord<ENC, 2, FMODE, 2>, mk::kScanThreads, 2 * 32};
        default: return {mk::mk_scan_ord<ENC, 1, FMODE, 2>, mk::kScanThreads, 2 * 32};
    }
}

// Stride 8, ASCII, L2-resident dual-key filter (large query sets). MK_DUAL_SHAPE picks another launch shape
// (tuning runs only).
ScanLaunch pick_dual8(bool gate) {
    // measured (cfg5, upper-case queries, scan ms): U=2 T=1024: 1.06, U=4 T=768: 1.08, U=4 T=512: 1.14
    const int shape = std::getenv("MK_DUAL_SHAPE") ? std::atoi(std::getenv("MK_DUAL_SHAPE")) : 1;
    switch (shape) {
        case 0: return gate ? ScanLaunch{mk::mk_scan_dual8<4, 512, true>, 512, 4 * 32} : ScanLaunch{mk::mk_scan_dual8<4, 512, false>, 512, 4 * 32};
        case 2: return gate ? ScanLaunch{mk::mk_scan_dual8<4, 768, true>, 768, 4 * 32} : ScanLaunch{mk::mk_scan_dual8<4, 768, false>, 768, 4 * 32};
        case 3: return gate ? ScanLaunch{mk::mk_scan_dual8<2, 512, true>, 512, 2 * 32} : ScanLaunch{mk::mk_scan_dual8<2, 512, false>, 512, 2 * 32};
        default: return gate ? ScanLaunch{mk::mk_scan_dual8<2, 1024, true>, 1024, 2 * 32} : ScanLaunch{mk::mk_scan_dual8<2, 1024, false>, 1024, 2 * 32};
    }
}

// MK_TMA=1|2|3 (experiment, see mk_scan_d16_tma): tiles staged by bulk copies; 1: U=2 T=896, 2: U=4 T=512, 3: U=2 T=512.
// The shapes are bounded by shared memory: filter (128-160 KiB) + candidate queues + 2 stages of U x 512 bytes per warp.
int tma_shape() {
    static const int v = std::getenv("MK_TMA") ? std::atoi(std::getenv("MK_TMA")) : 0;
    return v;
}
bool use_tma(const mk::Tables& t) { return tma_shape() > 0 && t.d == 16 && t.filter_in_smem; }
template <int ENC, bool B32>
ScanLaunch pick_tma() {
    switch (tma_shape()) {
        case 2: return {mk::mk_scan_d16_tma<ENC, 4, 512, B32>, 512, 4 * 32};
        case 3: return {mk::mk_scan_d16_tma<ENC, 2, 512, B32>, 512, 2 * 32};
        default: return {mk::mk_scan_d16_tma<ENC, 2, 896, B32>, 896, 2 * 32};
    }
}

// Strides 1 and 2, ASCII, direct prefix bitmap in shared memory (shortest pattern below 15 bases). MK_SHORT_SHAPE picks
// another launch shape (tuning runs only).
ScanLaunch pick_short(uint32_t d) {
    static const int shape = std::getenv("MK_SHORT_SHAPE") ? std::atoi(std::getenv("MK_SHORT_SHAPE")) : 0;
    if (d == 1) {
        switch (shape) {
            case 1: return {mk::mk_scan_short<1, 4, 512>, 512, 4 * 32};
            case 3: return {mk::mk_scan_short<1, 2, 1024>, 1024, 2 * 32};  // (spills at 64 registers)
            default: return {mk::mk_scan_short<1, 2, 768>, 768, 2 * 32};
        }
    }
    switch (shape) {
        case 1: return {mk::mk_scan_short<2, 4, 512>, 512, 4 * 32};
        case 2: return {mk::mk_scan_short<2, 2, 768>, 768, 2 * 32};
        default: return {mk::mk_scan_short<2, 2, 1024>, 1024, 2 * 32};
    }
}

// The scan kernel of a table set
ScanLaunch pick_kernel(int enc, uint32_t d, bool smemf, bool win = false, bool f32 = false);
ScanLaunch pick_kernel(const mk::Tables& t) {
    if (t.filter_direct) return pick_short(t.d);
    if (t.dual_perm) return pick_dual8(t.gate_mask != 0);
    if (use_tma(t)) {
        if (t.enc == MK_ENC_ASCII) return t.filter32 ? pick_tma<MK_ENC_ASCII, true>() : pick_tma<MK_ENC_ASCII, false>();
        return t.filter32 ? pick_tma<MK_ENC_BAM4, true>() : pick_tma<MK_ENC_BAM4, false>();
    }
    return pick_kernel(t.enc, t.d, t.filter_in_smem, t.win, t.filter32);
}
size_t scan_smem_bytes(const mk::Tables& t) {
    if (!t.filter_in_smem) return 0;
    if (use_tma(t)) {
        const ScanLaunch k = pick_kernel(t);
        return ((t.filter.size() * 4 + 127) & ~(size_t)127) + (size_t)(k.threads / 32) * 2 * (size_t)(k.tile_vecs * 16);
    }
    return t.filter.size() * 4;
}

ScanLaunch pick_kernel(int enc, uint32_t d, bool smemf, bool win, bool f32) {
    if (f32) {  // shared-memory filter with 32-bit blocks (small seed sets)
        if (win) {
            if (enc == MK_ENC_ASCII && d == 8) return {mk::mk_scan_win<MK_ENC_ASCII, 8, 4, 896, true>, 896, 4 * 32};
            if (enc == MK_ENC_ASCII) return {mk::mk_scan_win<MK_ENC_ASCII, 4, 4, 768, true>, 768, 4 * 32};
            return {mk::mk_scan_win<MK_ENC_BAM4, 8, 4, 768, true>, 768, 4 * 32};
        }
        if (enc == MK_ENC_ASCII) return {mk::mk_scan_d16<MK_ENC_ASCII, mk::kFilterSmem, 4, 896, false, true>, 896, 4 * 32};
        if (d16_shape() != 2) return {mk::mk_scan_d16<MK_ENC_BAM4, mk::kFilterSmem, 4, 768, false, true>, 768, 4 * 32};
        return {mk::mk_scan_d16<MK_ENC_BAM4, mk::kFilterSmem, 4, 896, false, true>, 896, 4 * 32};
    }
    if (win) {  // stride 8 / 4, shared-memory filter, window seeds in the permuted packing
        if (enc == MK_ENC_ASCII && d == 8) return {mk::mk_scan_win<MK_ENC_ASCII, 8, 4, 896>, 896, 4 * 32};
#ifdef MK_TUNE_BUILD
        if (enc == MK_ENC_ASCII) {
            switch (tune_env("MK_WIN_VARIANT", 0)) {
                case 1: return {mk::mk_scan_win<MK_ENC_ASCII, 4, 2, 1024>, 1024, 2 * 32};
                case 2: return {mk::mk_scan_win<MK_ENC_ASCII, 4, 4, 640>, 640, 4 * 32};
                case 3: return {mk::mk_scan_win<MK_ENC_ASCII, 4, 2, 896>, 896, 2 * 32};
                case 4: return {mk::mk_scan_win<MK_ENC_ASCII, 4, 4, 896>, 896, 4 * 32};
                default: break;
            }
        }
#endif
        if (enc == MK_ENC_ASCII) return {mk::mk_scan_win<MK_ENC_ASCII, 4, 4, 768>, 768, 4 * 32};
        return {mk::mk_scan_win<MK_ENC_BAM4, 8, 4, 768>, 768, 4 * 32};
    }
    if (enc == MK_ENC_ASCII)
        return smemf ? pick_by_d<MK_ENC_ASCII, mk::kFilterSmem>(d) : pick_by_d<MK_ENC_ASCII, mk::kFilterGlobal>(d);
    return smemf ? pick_by_d<MK_ENC_BAM4, mk::kFilterSmem>(d) : pick_by_d<MK_ENC_BAM4, mk::kFilterGlobal>(d);
}

int init_workspace(Workspace& ws) {
    CU(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&ws.ev_begin));
    CU(cudaEventCreate(&ws.ev_scan));
    CU(cudaEventCreate(&ws.ev_verify));
    CU(cudaEventCreate(&ws.ev_end));
    CU(ws.counters.ensure(8));
    CU(ws.h_counters.ensure(8));
    CU(ws.radix_table.ensure((size_t)256 * mk::kSortWarps));
    CU(ws.radix_totals.ensure(256));
    CU(ws.buckets.ensure(3 * mk::kBuckets + 16));
    CU(cudaMemset(ws.buckets.p, 0, (3 * mk::kBuckets + 16) * sizeof(uint32_t)));
    if (std::getenv("MK_NO_BUCKET_SORT")) ws.prefer_radix = true;
    return MK_OK;
}

int ensure_hit_capacity(Workspace& ws, uint64_t cap) {
    if (cap <= ws.hit_cap) return MK_OK;
    CU(ws.raw_a.ensure(cap));
    CU(ws.raw_b.ensure(cap));
    CU(cudaMemset(ws.raw_b.p, 0, cap * sizeof(mk::RawHit)));  // never holds anything but valid pattern ids (see enqueue_sort)
    CU(ws.out.ensure(cap));
    CU(ws.heads.ensure(cap / mk::kHeadTile + 2));
    ws.hit_cap = cap;
    return MK_OK;
}

int ensure_tables(mk_engine* e, int enc) {
    DeviceTables& dt = e->tables[enc];
    if (dt.built) return MK_OK;
    try {
        dt.host = e->tab->get(enc);
    } catch (const std::bad_alloc&) {
        return fail(MK_ERR_NOMEM, "out of host memory while building the seed tables");
    } catch (const std::exception& ex) {
        return fail(MK_ERR_INVALID, "table build failed: %s", ex.what());
    }
    CU(dt.filter.upload(dt.host->filter));
    if (!dt.host->filter2.empty()) CU(dt.filter2.upload(dt.host->filter2));
    CU(dt.postings.upload(dt.host->postings));
    CU(dt.pat_off.upload(dt.host->pat_off));
    CU(dt.slots.upload(dt.host->slots));
    CU(dt.pat_bytes.upload(dt.host->pat_bytes));
    ScanLaunch k = pick_kernel(*dt.host);
    if (scan_smem_bytes(*dt.host))
        CU(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes(*dt.host)));
    if (const char* co = std::getenv("MK_CARVEOUT"))  // tuning: shared-memory carve-out of the scan kernel in per cent
        CU(cudaFuncSetAttribute(k.fn, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(co)));
    dt.built = true;
    return MK_OK;
}

// Sort the raw hit list of ws (raw_a) into the report order and convert it (ws.out). buckets: the bucket
// sort (4 launches; raises counters[3] if a bucket overflows, see finish_batch), else the radix passes.
int enqueue_sort(mk_engine* e, Workspace& ws, bool buckets) {
    DeviceTables& dt = e->tables[ws.enc];
    const uint32_t key_bits = ws.key_bits;
    const uint32_t len_tie_bits = e->tab->ps.len_bits + e->tab->ps.tie_bits;
    const unsigned long long* cnt = ws.counters.p;
    ws.used_buckets = buckets;
    if (buckets) {
        // the largest key of the batch: ALL_HITS (end <= n_units | len field | tie), PATTERN_SET (record | pattern)
        const unsigned long long max_key = ws.mode == MK_MODE_ALL_HITS
                                               ? ((((unsigned long long)ws.n_units + 1) << len_tie_bits) - 1)
                                               : ((((unsigned long long)ws.n_records) << mk::bits_for(e->tab->ps.n ? e->tab->ps.n - 1 : 0)) - 1);
        const mk::BucketMap shift = mk::make_bucket_map(max_key);
        (void)key_bits;
        uint32_t* bcount = ws.buckets.p;
        uint32_t* bstart = ws.buckets.p + mk::kBuckets;
        uint32_t* bcursor = ws.buckets.p + 2 * mk::kBuckets + 8;
        unsigned long long* overflow = ws.counters.p + 3;
        mk::mk_bucket_count<<<e->sm_count * 2, 256, 0, ws.stream>>>(ws.raw_a.p, cnt, ws.hit_cap, shift, bcount);
        mk::mk_bucket_scan<<<1, 1024, 0, ws.stream>>>(bcount, bstart, bcursor, overflow);
        mk::mk_bucket_scatter<<<e->sm_count * 8, 256, 0, ws.stream>>>(ws.raw_a.p, ws.raw_b.p, cnt, ws.hit_cap, shift, bcursor, overflow);
        if (ws.mode == MK_MODE_ALL_HITS) {
            mk::mk_bucket_sort<<<e->sm_count * 7, 256, 0, ws.stream>>>(ws.raw_b.p, bstart, overflow, ws.out.p, ws.d_off, dt.pat_off.p, len_tie_bits);
        } else {
            mk::mk_bucket_sort<<<e->sm_count * 7, 256, 0, ws.stream>>>(ws.raw_b.p, bstart, overflow, nullptr, nullptr, nullptr, 0);
            mk::mk_heads_count<<<e->sm_count * 4, 256, 0, ws.stream>>>(ws.raw_b.p, cnt, ws.hit_cap, ws.heads.p);
            mk::mk_heads_scan<<<1, 1024, 0, ws.stream>>>(cnt, ws.hit_cap, ws.heads.p, ws.counters.p + 1);
            mk::mk_finalize_pairs<<<e->sm_count * 4, 256, 0, ws.stream>>>(ws.raw_b.p, ws.out.p, cnt, ws.hit_cap, ws.heads.p, dt.pat_off.p);
        }
        CU(cudaGetLastError());
        return MK_OK;
    }
    {
        mk::RawHit *src = ws.raw_a.p, *dst = ws.raw_b.p;
        for (uint32_t shift = 0; shift < key_bits; shift += 8) {
            mk::mk_radix_hist<<<mk::kSortBlocks, mk::kSortThreads, 0, ws.stream>>>(src, cnt, ws.hit_cap, shift, ws.radix_table.p);
            mk::mk_radix_rowscan<<<256, 256, 0, ws.stream>>>(cnt, ws.hit_cap, ws.radix_table.p, ws.radix_totals.p);
            mk::mk_radix_scatter<<<mk::kSortBlocks, mk::kSortThreads, 0, ws.stream>>>(src, dst, cnt, ws.hit_cap, shift,
                                                                                    ws.radix_table.p, ws.radix_totals.p);
            std::swap(src, dst);
        }
        if (ws.mode == MK_MODE_ALL_HITS) {
            mk::mk_finalize_hits<<<256, 256, 0, ws.stream>>>(src, ws.out.p, cnt, ws.hit_cap, ws.d_off, dt.pat_off.p, len_tie_bits);
        } else {
            mk::mk_heads_count<<<e->sm_count * 4, 256, 0, ws.stream>>>(src, cnt, ws.hit_cap, ws.heads.p);
            mk::mk_heads_scan<<<1, 1024, 0, ws.stream>>>(cnt, ws.hit_cap, ws.heads.p, ws.counters.p + 1);
            mk::mk_finalize_pairs<<<e->sm_count * 4, 256, 0, ws.stream>>>(src, ws.out.p, cnt, ws.hit_cap, ws.heads.p, dt.pat_off.p);
        }
        CU(cudaGetLastError());
    }
    return MK_OK;
}

// Candidate positions travel from the scan to the verify kernel as 32-bit grid indices (position / pos_mul):
// vector index * seeds per vector for the stride-16 scan, vector index * windows per vector for the window
// scans, the base position itself for the ordered scan. A batch whose last index does not fit is refused
// (MK_ERR_CAPACITY) instead of being verified at wrapped positions; INTEGRATION.md lists the limits.
int check_position_range(const mk::Tables& t, mk_encoding enc, uint64_t n_units) {
    const uint64_t seq_bytes = enc == MK_ENC_ASCII ? n_units : (n_units + 1) / 2;
    const uint64_t n_vec = (seq_bytes + 15) / 16;
    if (n_vec > 0xFFFFFFF0ull) return fail(MK_ERR_CAPACITY, "batch larger than 64 GiB of sequence");
    uint64_t per_vec;  // grid indices per 16-byte vector
    if (t.d == 16) per_vec = enc == MK_ENC_ASCII ? 1 : 2;
    else if (t.win) per_vec = (enc == MK_ENC_ASCII ? 16 : 32) / t.d;
    else return n_units >= (1ull << 32) ? fail(MK_ERR_CAPACITY, "batches of 2^32 bases or more need patterns of at least 31 bases; split the batch") : MK_OK;
    if (n_vec * per_vec >= (1ull << 32))
        return fail(MK_ERR_CAPACITY, "batch of %llu bases exceeds the 32-bit candidate index of this query set (seed stride %u: at most %llu bases per batch); split the batch",
                    (unsigned long long)n_units, t.d, (unsigned long long)(((1ull << 32) / per_vec - 1) * (enc == MK_ENC_ASCII ? 16 : 32)));
    return MK_OK;
}

// One candidate per thread; CTAs of 128 threads (MK_VERIFY_THREADS=256: 256; same speed when the kernel runs alone, and
// the smaller CTAs find room earlier while the last CTAs of a scan kernel on another stream are still running).
void launch_verify(mk_engine* e, Workspace& ws, const mk::ScanParams& P) {
    const int vt = std::getenv("MK_VERIFY_THREADS") ? std::atoi(std::getenv("MK_VERIFY_THREADS")) : 128;
    const int ctas = e->sm_count * 8 * 256 / (vt == 256 ? 256 : 128);
    if (vt == 256) {
        if (ws.enc == MK_ENC_ASCII) mk::mk_verify_candidates<MK_ENC_ASCII, 256><<<ctas, 256, 0, ws.stream>>>(P);
        else mk::mk_verify_candidates<MK_ENC_BAM4, 256><<<ctas, 256, 0, ws.stream>>>(P);
    } else {
        if (ws.enc == MK_ENC_ASCII) mk::mk_verify_candidates<MK_ENC_ASCII, 128><<<ctas, 128, 0, ws.stream>>>(P);
        else mk::mk_verify_candidates<MK_ENC_BAM4, 128><<<ctas, 128, 0, ws.stream>>>(P);
    }
}

// Enqueue the device work of one batch on ws.stream (no host synchronisation).
int enqueue(mk_engine* e, Workspace& ws) {
    DeviceTables& dt = e->tables[ws.enc];
    const mk::Tables& t = *dt.host;
    const uint64_t seq_bytes = ws.enc == MK_ENC_ASCII ? ws.n_units : (ws.n_units + 1) / 2;
    const size_t flag_words32 = ((size_t)ws.n_records + 63) / 64 * 2;

    mk::ScanParams P{};
    P.text = reinterpret_cast<const uint4*>(ws.d_seq);
    P.n_units = ws.n_units;
    P.n_vec = (uint32_t)((seq_bytes + 15) / 16);
    P.off = ws.d_off;
    P.lens = ws.d_lens;
    P.n_records = ws.n_records;
    P.filter = dt.filter.p;
    P.filter_log2_bits = t.filter_log2_bits;
    P.filter_blocks = t.filter_blocks;
    P.filter32 = t.filter32 ? 1u : 0u;
    P.filter2 = t.filter2.empty() ? nullptr : dt.filter2.p;
    P.filter2_log2_bits = t.filter2_log2_bits;
    P.slots = dt.slots.p;
    P.bucket_mask = t.bucket_mask;
    P.postings = dt.postings.p;
    P.pat_bytes = dt.pat_bytes.p;
    P.pat_off = dt.pat_off.p;
    P.tie_rank = e->tie_rank.p;
    P.q = t.q;
    P.short_shift = (t.perm || t.win || t.dual_perm) ? 0u : 32u - 2u * t.q;
    P.win_mask0 = t.win_mask0;
    P.win_mask1 = t.win_mask1;
    P.has_long = t.q2 ? 1u : 0u;
    P.case_insensitive = e->tab->ps.case_insensitive ? 1 : 0;
    P.short_mask = t.dual_perm ? t.win_mask0 : 0xFFFFFFFFu;
    P.pos_flags2 = t.dual_perm ? 1u : 0u;
    P.direct_shift = t.filter_direct ? 32u - 2u * t.direct_q1 : 0u;
    P.direct_words = t.filter_direct ? (uint32_t)t.filter.size() : 0u;
    P.gate_mask = t.gate_mask;
    P.gate_val = t.gate_val;
    P.cand = ws.cand.p;
    P.cand_capacity = ws.cand_cap;
    P.cand_count = ws.counters.p + 2;
    P.pos_mul = t.d == 16 ? MK_UNIT_BASES : (t.win ? t.d : 1);
    P.flags = ws.flags.p;
    P.hits = ws.raw_a.p;
    P.hit_capacity = ws.hit_cap;
    P.hit_count = ws.counters.p;
    P.mode = ws.mode;
    P.len_bits = e->tab->ps.len_bits;
    P.tie_bits = e->tab->ps.tie_bits;
    P.pat_bits = mk::bits_for(e->tab->ps.n ? e->tab->ps.n - 1 : 0);
    P.max_len = e->tab->ps.max_len;
    P.n_patterns = e->tab->ps.n;
    P.n_postings = (uint32_t)t.postings.size();
    P.cta_clock = nullptr;
    if (std::getenv("MK_CTA_CLOCKS")) {
        CU(ws.cta_clock.ensure(3 * (size_t)e->sm_count));
        CU(cudaMemsetAsync(ws.cta_clock.p, 0, 3 * (size_t)e->sm_count * sizeof(unsigned long long), ws.stream));
        P.cta_clock = ws.cta_clock.p;
    }

    uint32_t key_bits;
    if (ws.mode == MK_MODE_ALL_HITS) key_bits = mk::bits_for(ws.n_units) + P.len_bits + P.tie_bits;
    else key_bits = mk::bits_for(ws.n_records) + P.pat_bits;
    if (ws.mode != MK_MODE_FLAG && key_bits > 64)
        return fail(MK_ERR_CAPACITY, "batch too large for the 64-bit hit sort key (%u bits)", key_bits);

    int rc_pos = check_position_range(t, ws.enc, ws.n_units);
    if (rc_pos) return rc_pos;
    // device_ns is the device work of the batch: clearing the flags and counters belongs to it
    CU(cudaEventRecord(ws.ev_begin, ws.stream));
    CU(cudaMemsetAsync(ws.flags.p, 0, flag_words32 * 4, ws.stream));
    CU(cudaMemsetAsync(ws.counters.p, 0, 8 * sizeof(unsigned long long), ws.stream));
    if (P.n_vec > 0 && ws.n_records > 0) {
        ScanLaunch k = pick_kernel(t);
        const uint64_t warps = k.threads / 32;
        uint64_t tiles = ((uint64_t)P.n_vec + k.tile_vecs - 1) / k.tile_vecs;
        uint64_t want = (tiles + warps - 1) / warps;
        int grid = (int)std::min<uint64_t>((uint64_t)e->sm_count, std::max<uint64_t>(want, 1));
        k.fn<<<grid, k.threads, scan_smem_bytes(t), ws.stream>>>(P);
        CU(cudaEventRecord(ws.ev_scan, ws.stream));
        launch_verify(e, ws, P);
        CU(cudaGetLastError());
    } else {
        CU(cudaEventRecord(ws.ev_scan, ws.stream));
    }
    CU(cudaEventRecord(ws.ev_verify, ws.stream));
    ws.key_bits = key_bits;
    ws.used_buckets = false;
    if (ws.mode != MK_MODE_FLAG) {
        int rc = enqueue_sort(e, ws, !ws.prefer_radix);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ws.ev_end, ws.stream));
    CU(cudaMemcpyAsync(ws.h_counters.p, ws.counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ws.stream));
    if (ws.fetch)
        CU(cudaMemcpyAsync(ws.h_flags.p, ws.flags.p, flag_words32 * 4, cudaMemcpyDeviceToHost, ws.stream));
    return MK_OK;
}

int begin_batch(mk_engine* e, Workspace& ws, const void* d_seq, const unsigned long long* d_off, const uint32_t* d_lens,
                uint32_t n_records, uint64_t n_units, mk_encoding enc, mk_mode mode, bool fetch) {
    if (enc != MK_ENC_ASCII && enc != MK_ENC_BAM4) return fail(MK_ERR_INVALID, "unknown encoding %d", (int)enc);
    if (mode != MK_MODE_FLAG && mode != MK_MODE_PATTERN_SET && mode != MK_MODE_ALL_HITS)
        return fail(MK_ERR_INVALID, "unknown mode %d", (int)mode);
    int rc = ensure_tables(e, enc);
    if (rc) return rc;
    rc = check_position_range(*e->tables[enc].host, enc, n_units);  // before anything is sized for the batch
    if (rc) return rc;
    ws.d_seq = d_seq; ws.d_off = d_off; ws.d_lens = d_lens;
    ws.n_records = n_records; ws.n_units = n_units; ws.enc = enc; ws.mode = mode; ws.fetch = fetch;
    const size_t flag_words64 = ((size_t)n_records + 63) / 64;
    CU(ws.flags.ensure(std::max<size_t>(flag_words64 * 2, 2)));
    if (fetch) CU(ws.h_flags.ensure(std::max<size_t>(flag_words64, 1)));
    if (mode != MK_MODE_FLAG) {
        uint64_t cap = ws.hit_cap ? ws.hit_cap : (e->cfg.hit_capacity ? e->cfg.hit_capacity : (1ull << 20));
        rc = ensure_hit_capacity(ws, cap);
        if (rc) return rc;
    }
    {   // candidate list: grown on demand like the hit list
        uint64_t bytes = enc == MK_ENC_ASCII ? n_units : (n_units + 1) / 2;
        uint64_t want = std::max<uint64_t>(1ull << 20, bytes / 1024);
        if (ws.cand_cap < want) {
            CU(ws.cand.ensure(want));
            ws.cand_cap = want;
        }
    }
    rc = enqueue(e, ws);
    if (rc) return rc;
    ws.busy = true;
    return MK_OK;
}

int finish_batch(mk_engine* e, Workspace& ws, mk_result* out) {
    if (!ws.busy) return fail(MK_ERR_STATE, "no batch in flight");
    uint32_t rescans = 0;
    float ms_total = 0.f, ms_scan = 0.f, ms_verify = 0.f;
    for (;;) {
        CU(cudaStreamSynchronize(ws.stream));
        float a = 0.f, b = 0.f, c = 0.f;
        CU(cudaEventElapsedTime(&a, ws.ev_begin, ws.ev_end));
        CU(cudaEventElapsedTime(&b, ws.ev_begin, ws.ev_scan));
        CU(cudaEventElapsedTime(&c, ws.ev_scan, ws.ev_verify));
        ms_total += a; ms_scan += b; ms_verify += c;
        if (ws.cta_clock.p && std::getenv("MK_CTA_CLOCKS")) {  // diagnostics: when did the scan CTAs start and end?
            std::vector<unsigned long long> h(3 * (size_t)e->sm_count);
            CU(cudaMemcpy(h.data(), ws.cta_clock.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
            std::vector<unsigned long long> ends;
            for (int i = 0; i < e->sm_count; ++i) {
                if (!h[2 * i] || !h[2 * i + 1]) continue;
                s0 = std::min(s0, h[2 * i]); s1 = std::max(s1, h[2 * i]);
                e0 = std::min(e0, h[2 * i + 1]); e1 = std::max(e1, h[2 * i + 1]);
                ends.push_back(h[2 * i + 1]);
            }
            if (std::getenv("MK_CTA_CLOCKS")[0] == '2') {  // per CTA: SM id and duration
                std::fprintf(stderr, "[merkurio] scan CTA us by smid:");
                for (int i = 0; i < e->sm_count; ++i)
                    if (h[2 * i] && h[2 * i + 1]) std::fprintf(stderr, " %llu:%.0f", h[2 * (size_t)e->sm_count + i], (h[2 * i + 1] - h[2 * i]) / 1e3);
                std::fprintf(stderr, "\n");
            }
            if (!ends.empty()) {
                std::sort(ends.begin(), ends.end());
                std::fprintf(stderr, "[merkurio] scan CTAs: %zu, start spread %.1f us, kernel %.1f us, ends before the last CTA's: first %.1f, median %.1f, 90th percentile %.1f us\n",
                             ends.size(), (s1 - s0) / 1e3, (e1 - s0) / 1e3, (e1 - e0) / 1e3, (e1 - ends[ends.size() / 2]) / 1e3,
                             (e1 - ends[ends.size() * 9 / 10]) / 1e3);
            }
        }
        const bool cand_over = ws.h_counters.p[2] > ws.cand_cap;
        const bool hits_over = ws.mode != MK_MODE_FLAG && ws.h_counters.p[0] > ws.hit_cap;
        if (!cand_over && !hits_over && ws.used_buckets && ws.h_counters.p[3]) {
            // a bucket of the bucket sort overflowed (hits piled up in one key range): the raw list is untouched,
            // sort it with the radix passes, and keep to them on this workspace
            ws.prefer_radix = true;
            int rc = enqueue_sort(e, ws, false);
            if (rc) { ws.busy = false; return rc; }
            CU(cudaEventRecord(ws.ev_end, ws.stream));
            CU(cudaMemcpyAsync(ws.h_counters.p, ws.counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ws.stream));
            CU(cudaStreamSynchronize(ws.stream));
            float extra = 0.f;
            CU(cudaEventElapsedTime(&extra, ws.ev_verify, ws.ev_end));
            ms_total += extra;
        }
        if (!cand_over && !hits_over) break;
        // a list overflowed: grow it to the exact need and scan the batch again (never drop hits)
        int rc = MK_OK;
        if (cand_over) {
            uint64_t need = ws.h_counters.p[2] + ws.h_counters.p[2] / 8 + 1024;
            cudaError_t ce = ws.cand.ensure(need);
            if (ce != cudaSuccess) rc = fail(MK_ERR_NOMEM, "cannot grow the candidate list to %llu entries: %s", (unsigned long long)need, cudaGetErrorString(ce));
            else ws.cand_cap = need;
        }
        if (!rc && hits_over) {
            uint64_t need = ws.h_counters.p[0];
            rc = ensure_hit_capacity(ws, need + need / 8 + 1024);
        }
        if (rc) { ws.busy = false; return rc; }
        ++rescans;
        rc = enqueue(e, ws);
        if (rc) { ws.busy = false; return rc; }
    }
    ws.busy = false;
    uint64_t n_hits = 0;
    if (ws.mode == MK_MODE_ALL_HITS) n_hits = ws.h_counters.p[0];
    else if (ws.mode == MK_MODE_PATTERN_SET) n_hits = ws.h_counters.p[1];
    if (ws.fetch && n_hits) {
        CU(ws.h_hits.ensure(n_hits));
        CU(cudaMemcpyAsync(ws.h_hits.p, ws.out.p, n_hits * sizeof(mk_hit), cudaMemcpyDeviceToHost, ws.stream));
        CU(cudaStreamSynchronize(ws.stream));
    }
    if (out) {
        out->record_flags = ws.fetch ? ws.h_flags.p : nullptr;
        out->n_records = ws.n_records;
        out->reserved = 0;
        out->hits = (ws.fetch && n_hits) ? ws.h_hits.p : nullptr;
        out->n_hits = n_hits;
        out->bases_scanned = ws.n_units;
        out->device_ns = (uint64_t)((double)ms_total * 1e6);
        out->scan_ns = (uint64_t)((double)ms_scan * 1e6);
        out->verify_ns = (uint64_t)((double)ms_verify * 1e6);
        out->n_candidates = ws.h_counters.p[2];
        out->n_rescans = rescans;
        out->reserved2 = 0;

        out->d_record_flags = reinterpret_cast<const uint64_t*>(ws.flags.p);
        out->d_hits = n_hits ? ws.out.p : nullptr;
    }
    return MK_OK;
}

int get_slot(mk_engine* e, uint32_t slot, Slot** out) {
    if (!e) return fail(MK_ERR_INVALID, "null engine");
    if (slot >= e->slots.size()) return fail(MK_ERR_STATE, "slot %u out of range (engine has %zu)", slot, e->slots.size());
    *out = e->slots[slot].get();
    return MK_OK;
}

int check_batch(mk_engine* e, uint32_t n_records, uint64_t n_units, mk_encoding enc) {
    uint64_t bytes = enc == MK_ENC_ASCII ? n_units : (n_units + 1) / 2;
    if (bytes > e->cfg.max_batch_bytes)
        return fail(MK_ERR_CAPACITY, "batch of %llu bytes exceeds max_batch_bytes %llu", (unsigned long long)bytes,
                    (unsigned long long)e->cfg.max_batch_bytes);
    if (n_records > e->cfg.max_batch_records)
        return fail(MK_ERR_CAPACITY, "batch of %u records exceeds max_batch_records %u", n_records, e->cfg.max_batch_records);
    return MK_OK;
}

}  // namespace

extern "C" {

const char* mk_last_error(void) { return g_err.c_str(); }
const char* mk_version(void) {
#ifdef MK_DEBUG_CHECKS
    return "merkurio-b200 0.2.0 (sm_100a, device-side debug checks)";
#else
    return "merkurio-b200 0.2.0 (sm_100a)";
#endif
}

static int check_patterns(const mk_patterns* patterns) {
    if (patterns->n == 0) return fail(MK_ERR_NO_PATTERNS, "No k-mers found in file or provided sequence.");
    if (!patterns->bytes || !patterns->off) return fail(MK_ERR_INVALID, "null pattern storage");
    if (patterns->n > mk::kMaxPatternId) return fail(MK_ERR_INVALID, "too many patterns (%u)", patterns->n);
    for (uint32_t p = 0; p < patterns->n; ++p) {
        if (patterns->off[p + 1] < patterns->off[p]) return fail(MK_ERR_INVALID, "pattern offsets must be non-decreasing");
        if (patterns->off[p + 1] == patterns->off[p]) return fail(MK_ERR_EMPTY_PATTERN, "Pattern is empty.");
    }
    return MK_OK;
}

int mk_tables_create(const mk_patterns* patterns, int case_insensitive, mk_tables** out) {
    if (!patterns || !out) return fail(MK_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = check_patterns(patterns);
    if (rc) return rc;
    std::unique_ptr<mk_tables> t(new (std::nothrow) mk_tables);
    if (!t) return fail(MK_ERR_NOMEM, "out of memory");
    try {
        t->ps = mk::make_pattern_set(patterns->bytes, patterns->off, patterns->n, case_insensitive != 0);
    } catch (const std::bad_alloc&) {
        return fail(MK_ERR_NOMEM, "out of host memory");
    }
    // Large query sets take seconds to index (sort + cuckoo insertion): do it on host threads now, while the
    // caller creates CUDA contexts (also seconds on a multi-GPU box). Which encoding will be scanned is not
    // known yet, so both are prepared.
    if (patterns->n >= 50000 && !std::getenv("MK_NO_ASYNC_TABLES")) t->start_async();
    *out = t.release();
    return MK_OK;
}

void mk_tables_destroy(mk_tables* t) {
    if (t) t->release();
}

int mk_engine_create(const mk_patterns* patterns, const mk_config* config, mk_engine** out) {
    if (!patterns || !config || !out) return fail(MK_ERR_INVALID, "null argument");
    *out = nullptr;
    mk_tables* t = nullptr;
    int rc = mk_tables_create(patterns, config->case_insensitive, &t);
    if (rc) return rc;
    rc = mk_engine_create_shared(t, config, out);
    mk_tables_destroy(t);  // the engine holds its own reference
    return rc;
}

int mk_engine_create_shared(mk_tables* tables, const mk_config* config, mk_engine** out) {
    if (!tables || !config || !out) return fail(MK_ERR_INVALID, "null argument");
    *out = nullptr;
    if ((config->case_insensitive != 0) != tables->ps.case_insensitive)
        return fail(MK_ERR_INVALID, "mk_config.case_insensitive differs from the value the tables were created with");
    std::unique_ptr<mk_engine> e(new (std::nothrow) mk_engine);
    if (!e) return fail(MK_ERR_NOMEM, "out of memory");
    tables->refs.fetch_add(1);
    e->tab = tables;
    // MERKURIO_TIMING / MK_TIMING: where the start-up time goes (driver initialisation, context, allocations)
    const bool timing = std::getenv("MERKURIO_TIMING") || std::getenv("MK_TIMING");
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_c0 = now();
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    // On a box without the persistence daemon the GPU is re-initialised when the previous process has just
    // released the last context; a process starting at that moment can be turned away once. Anything but
    // "there is no device / no driver" is tried again a few times before it is reported.
    for (int attempt = 0; attempt < 4 && ce != cudaSuccess && ce != cudaErrorNoDevice && ce != cudaErrorInsufficientDriver; ++attempt) {
        cudaGetLastError();
        std::this_thread::sleep_for(std::chrono::milliseconds(250));
        ce = cudaGetDeviceCount(&ndev);
    }
    if (ce != cudaSuccess || ndev == 0)
        return fail(MK_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                    ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    if (config->device < 0 || config->device >= ndev) return fail(MK_ERR_INVALID, "device %d out of range", config->device);
    const double t_c1 = now();
    ce = cudaSetDevice(config->device);
    if (ce == cudaSuccess) ce = cudaFree(nullptr);  // the context is created here
    for (int attempt = 0; attempt < 4 && ce != cudaSuccess; ++attempt) {
        cudaGetLastError();
        std::this_thread::sleep_for(std::chrono::milliseconds(250));
        ce = cudaSetDevice(config->device);
        if (ce == cudaSuccess) ce = cudaFree(nullptr);
    }
    if (ce != cudaSuccess) return fail(MK_ERR_CUDA, "creating a CUDA context on device %d failed: %s", config->device, cudaGetErrorString(ce));
    const double t_c2 = now();
    e->device = config->device;
    e->cfg = *config;
    CU(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, e->device));
    CU(e->tie_rank.upload(e->tab->ps.tie_rank));
    int rc = init_workspace(e->direct);
    if (rc) return rc;
    const double t_c3 = now();
    for (uint32_t s = 0; s < config->n_slots; ++s) {
        std::unique_ptr<Slot> sl(new (std::nothrow) Slot);
        if (!sl) return fail(MK_ERR_NOMEM, "out of memory");
        rc = init_workspace(sl->ws);
        if (rc) return rc;
        size_t seq_cap = ((size_t)config->max_batch_bytes + 15) / 16 * 16 + 16;
        CU(sl->h_seq.ensure(seq_cap));
        CU(sl->d_seq.ensure(seq_cap));
        CU(sl->h_off.ensure((size_t)config->max_batch_records + 1));
        CU(sl->d_off.ensure((size_t)config->max_batch_records + 1));
        // the lens buffers are allocated on first use (only BAM batches carry explicit lengths)
        e->slots.push_back(std::move(sl));
    }
    if (timing)
        std::fprintf(stderr, "[merkurio] engine start-up on device %d: driver %.3f s, context %.3f s, workspace %.3f s, %u slots (pinned + device) %.3f s\n",
                     config->device, t_c1 - t_c0, t_c2 - t_c1, t_c3 - t_c2, config->n_slots, now() - t_c3);
    *out = e.release();
    return MK_OK;
}

void mk_engine_destroy(mk_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    delete e;
}

int mk_engine_get_info(mk_engine* e, mk_engine_info* out) {
    if (!e || !out) return fail(MK_ERR_INVALID, "null argument");
    *out = mk_engine_info{};
    out->n_patterns = e->tab->ps.n;
    out->min_len = e->tab->ps.min_len;
    out->max_len = e->tab->ps.max_len;
    out->sm_count = (uint32_t)e->sm_count;
    for (int enc = 0; enc < 2; ++enc)
        if (e->tables[enc].built) {
            const mk::Tables& t = *e->tables[enc].host;
            out->features |= ((t.dual_perm ? MK_FEATURE_DUAL8 : 0u) | (t.gate_mask ? MK_FEATURE_GATE : 0u)) << (8 * enc);
        }
    for (int enc = 0; enc < 2; ++enc) {
        const DeviceTables& dt = e->tables[enc];
        if (!dt.built) continue;
        out->seed_q[enc] = dt.host->q;
        out->seed_d[enc] = dt.host->d;
        out->n_seeds[enc] = dt.host->n_seeds;
        out->filter_log2_bits[enc] = dt.host->filter_in_smem ? 0 : dt.host->filter_log2_bits;
        out->filter_bytes[enc] = dt.host->filter.size() * 4;
        out->filter_hashes[enc] = dt.host->filter_hashes;
        out->filter_in_smem[enc] = dt.host->filter_in_smem ? 1 : 0;
        out->table_bytes[enc] = dt.bytes();
    }
    return MK_OK;
}

const char* mk_engine_scan_kernel(mk_engine* e, mk_encoding enc) {
    thread_local std::string name;
    name.clear();
    if (!e || (enc != MK_ENC_ASCII && enc != MK_ENC_BAM4) || !e->tables[enc].built) return "";
    const mk::Tables& t = *e->tables[enc].host;
    const char* en = enc == MK_ENC_ASCII ? "ASCII" : "BAM4";
    char buf[160];
    if (t.filter_direct)
        std::snprintf(buf, sizeof buf, "mk_scan_short<%s, stride %u, direct %u-base prefix bitmap>", en, t.d, t.direct_q1);
    else if (use_tma(t))
        std::snprintf(buf, sizeof buf, "mk_scan_d16_tma<%s, bulk-copy staged tiles, shape %d>", en, tma_shape());
    else if (t.dual_perm)
        std::snprintf(buf, sizeof buf, "mk_scan_dual8<%s, stride 8, L2 dual-key filter%s>", en, t.gate_mask ? ", alphabet gate" : "");
    else if (t.d == 16)
        std::snprintf(buf, sizeof buf, "mk_scan_d16<%s, %s>", en, t.filter_in_smem ? (t.filter32 ? "smem filter 32-bit blocks" : "smem filter 64-bit blocks") : "L2 bitmap");
    else if (t.win)
        std::snprintf(buf, sizeof buf, "mk_scan_win<%s, stride %u, smem filter %s>", en, t.d, t.filter32 ? "32-bit blocks" : "64-bit blocks");
    else
        std::snprintf(buf, sizeof buf, "mk_scan_ord<%s, stride %u, %s>", en, t.d, t.filter_in_smem ? "smem filter" : (t.filter_dual ? "L2 dual-key filter" : "L2 bitmap"));
    name = buf;
    return name.c_str();
}

int mk_slot_buffers(mk_engine* e, uint32_t slot, uint8_t** seq_pinned, uint64_t** off_pinned, uint32_t** lens_pinned) {
    Slot* s = nullptr;
    int rc = get_slot(e, slot, &s);
    if (rc) return rc;
    if (seq_pinned) *seq_pinned = s->h_seq.p;
    if (off_pinned) *off_pinned = s->h_off.p;
    if (lens_pinned) {
        CU(cudaSetDevice(e->device));
        CU(s->h_lens.ensure(std::max<size_t>(e->cfg.max_batch_records, 1)));
        *lens_pinned = s->h_lens.p;
    }
    return MK_OK;
}

int mk_scan_host(mk_engine* e, uint32_t slot, const uint8_t* h_seq, const uint64_t* h_off, const uint32_t* h_lens,
                 uint32_t n_records, uint64_t n_units, mk_encoding enc, mk_mode mode) {
    Slot* s = nullptr;
    int rc = get_slot(e, slot, &s);
    if (rc) return rc;
    if (s->ws.busy) return fail(MK_ERR_STATE, "slot %u still has a batch in flight", slot);
    if (!h_seq || !h_off) return fail(MK_ERR_INVALID, "null batch buffers");
    rc = check_batch(e, n_records, n_units, enc);
    if (rc) return rc;
    CU(cudaSetDevice(e->device));
    uint64_t bytes = enc == MK_ENC_ASCII ? n_units : (n_units + 1) / 2;
    cudaStream_t st = s->ws.stream;
    if (bytes) CU(cudaMemcpyAsync(s->d_seq.p, h_seq, bytes, cudaMemcpyHostToDevice, st));
    // the last vector is read whole: clear its tail so that the scan input is deterministic
    if (bytes % 16) CU(cudaMemsetAsync(s->d_seq.p + bytes, 0, 16 - bytes % 16, st));
    CU(cudaMemcpyAsync(s->d_off.p, h_off, ((size_t)n_records + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (h_lens) CU(s->d_lens.ensure(std::max<size_t>(e->cfg.max_batch_records, 1)));
    if (h_lens && n_records) CU(cudaMemcpyAsync(s->d_lens.p, h_lens, (size_t)n_records * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    return begin_batch(e, s->ws, s->d_seq.p, s->d_off.p, h_lens ? s->d_lens.p : nullptr, n_records, n_units, enc, mode, true);
}

namespace {
__global__ void mk_uniform_offsets(unsigned long long* off, uint32_t n_records, uint32_t record_len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_records; i += (uint64_t)gridDim.x * blockDim.x)
        off[i] = i * record_len;
}
}  // namespace

int mk_scan_host_uniform(mk_engine* e, uint32_t slot, const uint8_t* h_seq, uint32_t n_records, uint32_t record_len, mk_encoding enc,
                         mk_mode mode) {
    Slot* s = nullptr;
    int rc = get_slot(e, slot, &s);
    if (rc) return rc;
    if (s->ws.busy) return fail(MK_ERR_STATE, "slot %u still has a batch in flight", slot);
    if (!h_seq) return fail(MK_ERR_INVALID, "null batch buffer");
    if (enc == MK_ENC_BAM4 && (record_len & 1)) return fail(MK_ERR_INVALID, "BAM4 records of a common length must have an even length");
    const uint64_t n_units = (uint64_t)n_records * record_len;
    rc = check_batch(e, n_records, n_units, enc);
    if (rc) return rc;
    CU(cudaSetDevice(e->device));
    const uint64_t bytes = enc == MK_ENC_ASCII ? n_units : n_units / 2;
    cudaStream_t st = s->ws.stream;
    if (bytes) CU(cudaMemcpyAsync(s->d_seq.p, h_seq, bytes, cudaMemcpyHostToDevice, st));
    if (bytes % 16) CU(cudaMemsetAsync(s->d_seq.p + bytes, 0, 16 - bytes % 16, st));
    mk_uniform_offsets<<<std::max(1u, std::min(1024u, (n_records + 256) / 256)), 256, 0, st>>>(s->d_off.p, n_records, record_len);
    CU(cudaGetLastError());
    return begin_batch(e, s->ws, s->d_seq.p, s->d_off.p, nullptr, n_records, n_units, enc, mode, true);
}

int mk_scan_submit(mk_engine* e, uint32_t slot, uint32_t n_records, uint64_t n_units, int use_lens, mk_encoding enc,
                   mk_mode mode) {
    Slot* s = nullptr;
    int rc = get_slot(e, slot, &s);
    if (rc) return rc;
    if (use_lens && !s->h_lens.p) return fail(MK_ERR_STATE, "use_lens without lens_pinned from mk_slot_buffers");
    return mk_scan_host(e, slot, s->h_seq.p, s->h_off.p, use_lens ? s->h_lens.p : nullptr, n_records, n_units, enc, mode);
}

int mk_scan_wait(mk_engine* e, uint32_t slot, mk_result* out) {
    Slot* s = nullptr;
    int rc = get_slot(e, slot, &s);
    if (rc) return rc;
    CU(cudaSetDevice(e->device));
    return finish_batch(e, s->ws, out);
}

int mk_scan_device_submit(mk_engine* e, uint32_t slot, const void* d_seq, const uint64_t* d_off, const uint32_t* d_lens,
                          uint32_t n_records, uint64_t n_units, mk_encoding enc, mk_mode mode, int fetch) {
    Slot* s = nullptr;
    int rc = get_slot(e, slot, &s);
    if (rc) return rc;
    if (s->ws.busy) return fail(MK_ERR_STATE, "slot %u still has a batch in flight", slot);
    if ((!d_seq && n_units) || !d_off) return fail(MK_ERR_INVALID, "null device buffers");
    if (reinterpret_cast<uintptr_t>(d_seq) % 16) return fail(MK_ERR_INVALID, "d_seq must be 16-byte aligned");
    CU(cudaSetDevice(e->device));
    // Device-resident batches run in submission order: they would only compete for the same SMs, and the per-batch
    // CUDA-event times stay those of the batch alone. Measured alternatives (scripts/bench_overlap.py, DESIGN.md section 4,
    // profiles/r2_overlap_experiment/): the slots' streams left to run freely (MK_FREE=1), or all scans on one
    // high-priority stream with the post-processing of the batch before beside them — 2-12 % more throughput for a
    // stream of stride-16 batches, a loss of 17-60 % for the L2-filter scans, and no clean per-kernel times.
    if (!std::getenv("MK_FREE") && e->last_device_submit && e->last_device_submit != s->ws.ev_end)
        CU(cudaStreamWaitEvent(s->ws.stream, e->last_device_submit, 0));
    rc = begin_batch(e, s->ws, d_seq, reinterpret_cast<const unsigned long long*>(d_off), d_lens, n_records, n_units, enc, mode,
                     fetch != 0);
    if (rc == MK_OK) e->last_device_submit = s->ws.ev_end;
    return rc;
}

int mk_scan_device(mk_engine* e, const void* d_seq, const uint64_t* d_off, const uint32_t* d_lens, uint32_t n_records,
                   uint64_t n_units, mk_encoding enc, mk_mode mode, int fetch, mk_result* out) {
    if (!e) return fail(MK_ERR_INVALID, "null engine");
    if ((!d_seq && n_units) || !d_off) return fail(MK_ERR_INVALID, "null device buffers");
    if (reinterpret_cast<uintptr_t>(d_seq) % 16) return fail(MK_ERR_INVALID, "d_seq must be 16-byte aligned");
    CU(cudaSetDevice(e->device));
    int rc = begin_batch(e, e->direct, d_seq, reinterpret_cast<const unsigned long long*>(d_off), d_lens, n_records, n_units,
                         enc, mode, fetch != 0);
    if (rc) return rc;
    return finish_batch(e, e->direct, out);
}

}  // extern "C"
