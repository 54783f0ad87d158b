// One gzip member decoded by several threads (the reference reads .gz inputs through needletail / flate2 on one
// thread; README.md:39). DEFLATE has no index and every match may reach 32 KiB back, so a stream cannot simply be cut
// into pieces. What this file does (the scheme of pugz and rapidgzip, written for this host's decoder):
//
//   * the compressed file is mapped and cut at fixed byte offsets (2 MiB); each piece is a task for a pool of threads;
//   * a task looks for the first position in its piece where the header of a dynamic, non-final block stands
//     (Inflater::probe_dynamic_header: block type, code counts, a complete code-length code, complete literal /
//     distance codes — random bits do not pass), and decodes from there, block after block, up to the first block
//     boundary behind the end of its piece WITHOUT the 32 KiB of history: its output is 16-bit, and a match that
//     reaches in front of the start copies place holders ("byte j of the unknown window") instead of bytes;
//   * a stitcher thread walks the true sequence of blocks. It knows the exact bit position where the data decoded so
//     far ends. If a finished task started exactly there, its output is the continuation: the stitcher resolves the
//     task's last 32 KiB against the current window (that is the next window), hands the replacement of the place
//     holders in the rest to the pool, and jumps to the end of that task. If not (the boundary was a stored / fixed /
//     final block the search does not accept, or the search found something that is no boundary), it decodes one block
//     itself with the ordinary byte decoder and looks again. So a task's output is only ever used if the real decode
//     arrives at its first bit at a block boundary, where decoding with an unknown window is exact; a wrong guess of
//     the search costs time, never correctness;
//   * read() hands out the pieces in order, combining their CRC-32s.
//
// CRC-32 and length of every member are checked as in the sequential reader, members may follow each other, and every
// error of the sequential reader is raised at the same place with the same text (the sequential decoder is what runs
// wherever the tasks do not line up — at the latest at the damaged spot), after the data in front of it.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "codecs.h"
#include "common.h"
#include "inflate.h"

namespace mkh {

namespace {

constexpr size_t kWin = 32768;
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Task {  // one piece of the compressed file, decoded with an unknown window
    bool done = false;
    bool ok = false;           // a start was found and the data decoded up to end_bit
    bool ended_final = false;
    uint64_t start_bit = 0, end_bit = 0;
    std::vector<uint16_t> sym;  // kWin place holders + the decoded symbols (n_sym entries in use)
    size_t n_sym = 0;
};

struct Piece {  // one stretch of output, in the order of the stream
    enum Kind { kData, kMemberEnd, kError, kEnd } kind = kData;
    bool ready = false;
    std::vector<uint8_t> bytes;    // (a buffer that goes round: n_bytes of it are the data)
    size_t n_bytes = 0;
    uint32_t crc = 0;              // kData: of the data; kMemberEnd: the trailer's
    uint32_t isize = 0;            // kMemberEnd: the trailer's
    std::string error;             // kError
    // kData that still has place holders to replace (a job for the pool):
    std::vector<uint16_t> sym;
    size_t n_sym = 0;
    std::vector<uint8_t> window;   // the 32 KiB in front of it
    size_t window_valid = 0;
    bool bad_distance = false;
};

class ParallelGzipStream : public InputStream {
public:
    // The file, and one zero page behind it (the decoders read up to 600 bytes past the end of the data); nullptr if the
    // file cannot be mapped (the caller then reads it with the sequential reader).
    static const uint8_t* map_file(int fd, size_t size, size_t* map_len) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        *map_len = (size + page - 1) / page * page + page;
        void* m = ::mmap(nullptr, *map_len, PROT_READ, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) return nullptr;
        if (::mmap(m, size, PROT_READ, MAP_PRIVATE | MAP_FIXED, fd, 0) == MAP_FAILED) {
            ::munmap(m, *map_len);
            return nullptr;
        }
        ::madvise(m, size, MADV_SEQUENTIAL);
        return static_cast<const uint8_t*>(m);
    }

    ParallelGzipStream(int fd, const uint8_t* mapped, size_t map_len, size_t file_size, int threads, size_t piece_bytes)
        : fd_(fd), size_(file_size), piece_(piece_bytes), map_len_(map_len), base_(mapped), window_(kWin),
          seqbuf_(kWin + kSeqArea + Inflater::kOutputMargin + 64) {
        end_ = base_ + size_;
        timing_ = std::getenv("MERKURIO_TIMING") != nullptr;
        max_pieces_ = (size_t)threads + 2;
        lookahead_ = (size_t)threads + 2;
        size_t hdr = 0;
        const int hrc = parse_gzip_header(base_, size_, &hdr);
        if (hrc != 1) {  // (the caller has seen the magic bytes: a broken or cut header)
            push_error(hrc == 0 ? "Error while decompressing the input (truncated gzip stream)" : "Error while decompressing the input (gzip)");
            started_ = true;  // (nothing to start: the reader finds the error)
            return;
        }
        pos_ = (uint64_t)hdr * 8;
        data_begin_ = hdr;
        tasks_.resize(std::max<size_t>((size_ - hdr + piece_ - 1) / piece_, 1));
        n_threads_ = threads;  // (started by the first read that goes beyond the head of the data: see read())
    }
    ~ParallelGzipStream() override {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        cv_done_.notify_all();
        cv_piece_.notify_all();
        if (stitcher_.joinable()) stitcher_.join();
        for (auto& t : pool_) t.join();
        if (timing_ && !pool_.empty()) {
            std::fprintf(stderr, "[merkurio] gzip on %zu threads: %zu pieces of the file, %zu continued the decode where it stood, %zu not used; %zu blocks by the sequential decoder\n",
                         pool_.size(), tasks_.size(), n_used_, n_dropped_, n_seq_blocks_);
            std::fprintf(stderr, "[merkurio] gzip stitcher: waited %.3f s for tasks, %.3f s for the reader; pool: %.3f s searching, %.3f s decoding, %.3f s replacing place holders (sums over threads)\n",
                         t_wait_tasks_, t_wait_reader_, t_search_, t_decode_, t_resolve_);
        }
        ::munmap(const_cast<uint8_t*>(base_), map_len_);
        ::close(fd_);
    }

    size_t read(char* dst, size_t n) override {
        if (n == 0) return 0;
        if (!started_) {
            // The first 64 KiB come from a sequential decode of the head of the file, and the threads are only started
            // by a read that goes beyond it: the callers open every input a few times just to look at its first bytes
            // (is it there? FASTA or FASTQ?), and a pool that starts on a dozen pieces for each of those looks costs
            // tens of milliseconds of work that is thrown away.
            if (!head_tried_) {
                head_tried_ = true;
                head_.resize(kHead + Inflater::kOutputMargin + 64);
                Inflater h;
                const uint8_t* ip = base_ + data_begin_;
                uint8_t* op = head_.data();
                const Inflater::Status rc = h.run(&ip, end_, true, head_.data(), &op, head_.data() + head_.size() - 8);
                head_len_ = rc == Inflater::kError ? 0 : (size_t)(op - head_.data());  // (an error: reported by the full decode, in order)
            }
            if (head_pos_ < head_len_) {
                const size_t k = std::min(n, head_len_ - head_pos_);
                std::memcpy(dst, head_.data() + head_pos_, k);
                head_pos_ += k;
                return k;
            }
            std::vector<uint8_t>().swap(head_);
            skip_ = head_len_;  // the full decode starts at the first byte again: what was handed out already is dropped
            started_ = true;
            for (int t = 0; t < n_threads_; ++t) pool_.emplace_back([this] { worker(); });
            stitcher_ = std::thread([this] { stitch(); });
        }
        for (;;) {
            if (cur_ && cur_pos_ < cur_->n_bytes) {
                const size_t k = std::min(n, cur_->n_bytes - cur_pos_);
                std::memcpy(dst, cur_->bytes.data() + cur_pos_, k);
                cur_pos_ += k;
                return k;
            }
            if (ended_) {
                if (!error_.empty()) throw Error(error_);
                return 0;
            }
            // the next piece, once it is ready
            std::shared_ptr<Piece> p;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_piece_.wait(lk, [&] { return !pieces_.empty() && pieces_.front()->ready; });
                p = pieces_.front();
                pieces_.pop_front();
            }
            cv_piece_.notify_all();  // (room for the stitcher)
            if (cur_) give_bytes(std::move(cur_->bytes));
            cur_.reset();
            cur_pos_ = 0;
            switch (p->kind) {
                case Piece::kData:
                    if (p->bad_distance) fail("Error while decompressing the input (gzip)");
                    crc_ = p->n_bytes == 0 ? crc_ : (uint32_t)crc32_combine(crc_, p->crc, (z_off_t)p->n_bytes);
                    isize_ += (uint32_t)p->n_bytes;
                    cur_ = p;
                    if (skip_) {
                        cur_pos_ = std::min(skip_, p->n_bytes);
                        skip_ -= cur_pos_;
                    }
                    break;
                case Piece::kMemberEnd:
                    if (p->crc != crc_ || p->isize != isize_) fail("Error while decompressing the input (gzip)");
                    crc_ = 0;
                    isize_ = 0;
                    break;
                case Piece::kError:
                    fail(p->error);
                    break;
                case Piece::kEnd:
                    ended_ = true;
                    return 0;
            }
        }
    }

private:
    static constexpr size_t kSeqArea = 4u << 20, kHead = 64u << 10;

    [[noreturn]] void fail(const std::string& msg) {
        ended_ = true;
        error_ = msg;
        throw Error(msg);
    }

    // Buffers go round (tasks' symbols, pieces' bytes): a fresh 30 MB vector per task is thousands of page faults, and
    // page faults of a dozen threads hold up the CUDA start-up that runs beside the first seconds of a decode
    // (both want the process's address-space lock).
    std::vector<uint16_t> take_sym() {
        std::lock_guard<std::mutex> lk(mu_);
        if (free_sym_.empty()) return {};
        std::vector<uint16_t> v = std::move(free_sym_.back());
        free_sym_.pop_back();
        return v;
    }
    std::vector<uint8_t> take_bytes() {
        std::lock_guard<std::mutex> lk(mu_);
        if (free_bytes_.empty()) return {};
        std::vector<uint8_t> v = std::move(free_bytes_.back());
        free_bytes_.pop_back();
        return v;
    }
    void give_sym_locked(std::vector<uint16_t>&& v) {
        if (v.capacity() && free_sym_.size() < lookahead_ + max_pieces_) free_sym_.push_back(std::move(v));
        else std::vector<uint16_t>().swap(v);
    }
    void give_bytes(std::vector<uint8_t>&& v) {
        std::lock_guard<std::mutex> lk(mu_);
        if (v.capacity() && free_bytes_.size() < max_pieces_ + 2) free_bytes_.push_back(std::move(v));
    }

    static void huge_pages(void* p, size_t bytes) {
        const uintptr_t lo = ((uintptr_t)p + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1), hi = ((uintptr_t)p + bytes) & ~(((uintptr_t)2 << 20) - 1);
        if (hi > lo) ::madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
    }

    // ---- the pool: replacing place holders first (the reader waits for those), decoding pieces of the file otherwise --------
    void worker() {
        Inflater inf;
        for (;;) {
            std::shared_ptr<Piece> job;
            size_t i = 0;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || !jobs_.empty() || (next_task_ < tasks_.size() && next_task_ < consumed_ + lookahead_); });
                if (stop_) return;
                if (!jobs_.empty()) {
                    job = jobs_.front();
                    jobs_.pop_front();
                } else {
                    i = next_task_++;
                }
            }
            if (job) {
                const double t0 = now_s();
                try {
                    const size_t n = job->n_sym - kWin;
                    job->bytes = take_bytes();
                    if (job->bytes.capacity() < n) {
                        job->bytes.reserve(n + n / 4);
                        huge_pages(job->bytes.data(), job->bytes.capacity());
                    }
                    if (job->bytes.size() < n) job->bytes.resize(n);
                    job->n_bytes = n;
                    job->bad_distance = !resolve_markers(job->sym.data() + kWin, n, job->window.data(), job->window_valid, job->bytes.data());
                    job->crc = crc32_fast(0, job->bytes.data(), n);
                } catch (const std::exception&) {  // (no memory for the piece's bytes)
                    job->kind = Piece::kError;
                    job->error = "Error while decompressing the input (out of memory)";
                }
                std::vector<uint8_t>().swap(job->window);
                const double dt = now_s() - t0;
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    give_sym_locked(std::move(job->sym));
                    job->ready = true;
                    t_resolve_ += dt;
                }
                cv_piece_.notify_all();
                continue;
            }
            Task& t = tasks_[i];
            double t_search = 0, t_decode = 0;
            try {
                t.sym = take_sym();
                run_task(inf, i, t, &t_search, &t_decode);
            } catch (const std::exception&) {  // (no memory for the symbols: this piece is left to the sequential decoder)
                t.ok = false;
                std::vector<uint16_t>().swap(t.sym);
            }
            {
                std::lock_guard<std::mutex> lk(mu_);
                t.done = true;
                t_search_ += t_search;
                t_decode_ += t_decode;
            }
            cv_done_.notify_all();
        }
    }
    void run_task(Inflater& inf, size_t i, Task& t, double* t_search, double* t_decode) {
        const uint64_t lo = (uint64_t)(data_begin_ + i * piece_) * 8;
        const uint64_t hi = std::min<uint64_t>((uint64_t)(data_begin_ + (i + 1) * piece_), size_) * 8;
        uint64_t start = lo;
        const double t0 = now_s();
        if (i > 0) {  // (the first piece starts with the first block of the member, whatever its type)
            bool found = false;
            for (uint64_t p = lo; p < hi; ++p)
                if (inf.probe_dynamic_header(base_, end_, p)) { start = p; found = true; break; }
            if (!found) { *t_search = now_s() - t0; return; }
        }
        const double t1 = now_s();
        const uint64_t stop = (i + 1 == tasks_.size()) ? ~0ull : hi;
        // at most 24 Mi symbols per task (12 times the piece; the buffer doubles, so 64 MB at most): data that expands
        // further than that is left to the sequential decoder
        // (sized for a five-fold expansion at once: growing by doubling copies the symbols five times over)
        if (t.sym.capacity() < kWin + 5 * piece_ + 4096) {
            t.sym.reserve(kWin + 5 * piece_ + 4096);
            huge_pages(t.sym.data(), t.sym.capacity() * sizeof(uint16_t));  // (a sixth of the page faults where THP is on "madvise")
        }
        if (t.sym.size() < kWin + 5 * piece_ + 4096) t.sym.resize(kWin + 5 * piece_ + 4096);
        Inflater::MarkerRun r = inf.run_markers(base_, end_, start, stop, &t.sym, (size_t)12 * piece_);
        *t_search = t1 - t0;
        *t_decode = now_s() - t1;
        t.ok = r.ok;
        t.n_sym = r.n_out;
        t.ended_final = r.ended_final;
        t.start_bit = start;
        t.end_bit = r.end_bit;
    }

    // ---- the stitcher ----------------------------------------------------------------------------------------------
    // the oldest task that was not looked at yet, finished (blocks); nullptr when there is none left or at shutdown
    Task* oldest_task() {
        std::unique_lock<std::mutex> lk(mu_);
        if (consumed_ >= tasks_.size()) return nullptr;
        const double t0 = now_s();
        cv_done_.wait(lk, [&] { return stop_ || tasks_[consumed_].done; });
        t_wait_tasks_ += now_s() - t0;
        return stop_ ? nullptr : &tasks_[consumed_];
    }
    void drop_oldest_task() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            give_sym_locked(std::move(tasks_[consumed_].sym));
            tasks_[consumed_].sym = std::vector<uint16_t>();
            ++consumed_;
        }
        cv_work_.notify_all();
    }
    // appends a piece for the reader (waits while too many are queued); false at shutdown
    bool push_piece(std::shared_ptr<Piece> p, bool as_job) {
        {
            std::unique_lock<std::mutex> lk(mu_);
            const double t0 = now_s();
            cv_piece_.wait(lk, [&] { return stop_ || pieces_.size() < max_pieces_; });
            t_wait_reader_ += now_s() - t0;
            if (stop_) return false;
            pieces_.push_back(p);
            if (as_job) jobs_.push_back(p);
        }
        if (as_job) cv_work_.notify_all();
        else cv_piece_.notify_all();
        return true;
    }
    void push_error(const std::string& msg) {
        auto p = std::make_shared<Piece>();
        p->kind = Piece::kError;
        p->error = msg;
        p->ready = true;
        {
            std::lock_guard<std::mutex> lk(mu_);
            pieces_.push_back(p);
        }
        cv_piece_.notify_all();
    }
    // the last 32 KiB of the member's output after n more bytes p[0..n)
    void slide_window(const uint8_t* p, size_t n) {
        if (n >= kWin) {
            std::memcpy(window_.data(), p + n - kWin, kWin);
            window_valid_ = kWin;
        } else if (n) {
            std::memmove(window_.data(), window_.data() + n, kWin - n);
            std::memcpy(window_.data() + kWin - n, p, n);
            window_valid_ = std::min(kWin, window_valid_ + n);
        }
    }
    bool stopping() {
        std::lock_guard<std::mutex> lk(mu_);
        return stop_;
    }

    void stitch() {
        try {
            for (;;) {
                if (stopping()) return;
                if (member_done_) {
                    if (!next_member()) return;
                    continue;
                }
                bool took_task = false;
                while (!seq_active_) {  // (inside a block the sequential decoder goes on: tasks line up at block boundaries only)
                    Task* t = oldest_task();
                    if (!t) break;
                    if (!t->ok || t->start_bit < pos_) {  // found nothing, failed, or the real decode has passed its start already
                        ++n_dropped_;
                        drop_oldest_task();
                        continue;
                    }
                    if (t->start_bit > pos_) break;  // not there yet: one block by the sequential decoder below
                    // the real decode stands at the first bit of this task: its symbols are the continuation
                    auto p = std::make_shared<Piece>();
                    p->sym.swap(t->sym);
                    p->n_sym = t->n_sym;
                    p->window = window_;
                    p->window_valid = window_valid_;
                    const size_t n = p->n_sym - kWin;
                    // the next window: the last 32 KiB of this piece, resolved here (the rest is a job for the pool; a
                    // distance that reaches in front of the member shows there as well and is reported in order)
                    const size_t tail = std::min(n, kWin);
                    std::vector<uint8_t> tail_bytes(tail);
                    resolve_markers(p->sym.data() + kWin + n - tail, tail, window_.data(), window_valid_, tail_bytes.data());
                    slide_window(tail_bytes.data(), tail);
                    pos_ = t->end_bit;
                    if (t->ended_final) member_done_ = true;
                    ++n_used_;
                    drop_oldest_task();
                    if (n && !push_piece(p, true)) return;
                    took_task = true;
                    break;
                }
                if (took_task) continue;
                if (stopping()) return;
                if (!sequential_block()) return;
            }
        } catch (const Error& e) {
            push_error(e.what());
        } catch (const std::exception& e) {
            push_error(std::string("Error while decompressing the input (") + e.what() + ")");
        }
    }

    // one block (or as much of it as fits the buffer) from pos_ with the byte decoder, the history in front of it
    bool sequential_block() {
        if (!seq_active_) {
            seq_in_ = seq_.seek_bits(base_, pos_);
            seq_.set_stop_at_block_end(true);
            seq_active_ = true;
        }
        uint8_t* const area = seqbuf_.data() + kWin;
        std::memcpy(area - window_valid_, window_.data() + kWin - window_valid_, window_valid_);
        uint8_t* op = area;
        const Inflater::Status rc = seq_.run(&seq_in_, end_, true, area - window_valid_, &op, area + kSeqArea + Inflater::kOutputMargin);
        const size_t n = (size_t)(op - area);
        if (n) {
            auto p = std::make_shared<Piece>();
            p->bytes = take_bytes();
            if (p->bytes.size() < n) p->bytes.resize(n);
            std::memcpy(p->bytes.data(), area, n);
            p->n_bytes = n;
            p->crc = crc32_fast(0, area, n);
            p->ready = true;
            slide_window(area, n);
            if (!push_piece(p, false)) return false;
        }
        if (rc == Inflater::kError)
            throw Error(seq_in_ + 16 >= end_ ? "Error while decompressing the input (truncated gzip stream)" : "Error while decompressing the input (gzip)");
        if (rc == Inflater::kStreamEnd) member_done_ = true;
        if (rc == Inflater::kBlockEnd || rc == Inflater::kStreamEnd) {  // (kOutputFull: the block goes on with the next call)
            ++n_seq_blocks_;
            seq_active_ = false;
            pos_ = seq_.tell_bits(base_, seq_in_);
        }
        return true;
    }

    // trailer of the member that just ended, header of the next one if there is one; false when the stitcher is done
    bool next_member() {
        const size_t at = (size_t)((pos_ + 7) / 8);
        if (at + 8 > size_) throw Error("Error while decompressing the input (truncated gzip stream)");
        const uint8_t* t = base_ + at;
        auto p = std::make_shared<Piece>();
        p->kind = Piece::kMemberEnd;
        p->crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        p->isize = t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        p->ready = true;
        if (!push_piece(p, false)) return false;
        const size_t next = at + 8;
        if (next == size_) {
            auto e = std::make_shared<Piece>();
            e->kind = Piece::kEnd;
            e->ready = true;
            push_piece(e, false);
            return false;
        }
        size_t hdr = 0;
        const int rc = parse_gzip_header(base_ + next, size_ - next, &hdr);
        if (rc == 0) throw Error("Error while decompressing the input (truncated gzip stream)");
        if (rc < 0) throw Error("Error while decompressing the input (gzip)");
        pos_ = (uint64_t)(next + hdr) * 8;
        window_valid_ = 0;
        member_done_ = false;
        seq_active_ = false;
        return true;
    }

    int fd_;
    size_t size_, piece_;
    size_t map_len_ = 0;
    const uint8_t* base_ = nullptr;
    const uint8_t* end_ = nullptr;
    size_t data_begin_ = 0;
    bool timing_ = false;

    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_, cv_piece_;
    std::vector<Task> tasks_;
    std::deque<std::shared_ptr<Piece>> pieces_, jobs_;  // pieces_: in stream order, for the reader; jobs_: place holders to replace
    std::vector<std::vector<uint16_t>> free_sym_;
    std::vector<std::vector<uint8_t>> free_bytes_;
    std::vector<std::thread> pool_;
    std::thread stitcher_;
    size_t next_task_ = 0, consumed_ = 0, lookahead_ = 4, max_pieces_ = 4;
    bool stop_ = false;

    // stitcher state
    uint64_t pos_ = 0;  // bit position of the next block header of the real decode
    std::vector<uint8_t> window_;
    size_t window_valid_ = 0;
    std::vector<uint8_t> seqbuf_;
    Inflater seq_;
    const uint8_t* seq_in_ = nullptr;
    bool seq_active_ = false, member_done_ = false;
    size_t n_used_ = 0, n_dropped_ = 0, n_seq_blocks_ = 0;  // statistics (MERKURIO_TIMING)
    double t_wait_tasks_ = 0, t_wait_reader_ = 0, t_search_ = 0, t_decode_ = 0, t_resolve_ = 0;

    // reader state
    bool started_ = false, head_tried_ = false;
    int n_threads_ = 0;
    std::vector<uint8_t> head_;
    size_t head_len_ = 0, head_pos_ = 0, skip_ = 0;
    std::shared_ptr<Piece> cur_;
    size_t cur_pos_ = 0;
    bool ended_ = false;
    std::string error_;
    uint32_t crc_ = 0, isize_ = 0;
};

}  // namespace

// A gzip file decoded by several threads, or nullptr if that does not apply: not a regular file, too small to be worth
// it, or switched off (MERKURIO_GZIP_THREADS=1). Takes over fd when it returns a stream.
std::unique_ptr<InputStream> open_parallel_gzip(int fd) {
    struct stat st;
    if (::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return nullptr;
    const unsigned hw = std::thread::hardware_concurrency();
    int threads = (int)std::min(16u, std::max(2u, hw > 2 ? hw - 2 : 2u));  // (the stages behind the decoder mostly wait for it)
    if (const char* e = std::getenv("MERKURIO_GZIP_THREADS")) threads = std::atoi(e);
    if (threads < 2) return nullptr;
    size_t piece = (size_t)2 << 20;
    if (const char* e = std::getenv("MERKURIO_GZIP_PIECE_KB")) piece = std::max<size_t>(4, (size_t)std::atoll(e)) << 10;  // (tests: many small pieces)
    if ((size_t)st.st_size < 3 * piece) return nullptr;
    size_t map_len = 0;
    const uint8_t* mapped = ParallelGzipStream::map_file(fd, (size_t)st.st_size, &map_len);
    if (!mapped) return nullptr;  // (a file system without mmap: the sequential reader)
    return std::unique_ptr<InputStream>(new ParallelGzipStream(fd, mapped, map_len, (size_t)st.st_size, threads, piece));
}

}  // namespace mkh
