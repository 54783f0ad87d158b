// Block-parallel BGZF decompression (BAM, bgzip'ed FASTQ): BGZF is a series of independent gzip
// members of at most 64 KiB, each carrying its compressed size in a 'BC' extra field and its
// uncompressed size in the trailer, so the blocks of a stretch of the file can be inflated
// concurrently, each straight into its final place. Plays the role of the decompression threads the
// `bam` crate starts for the reference's `-p` option (src/cmd_tag.rs:504-507).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mkh {

// True if the file starts with a BGZF block header.
bool is_bgzf(int fd);

class BgzfReader {
public:
    // Takes ownership of fd (positioned anywhere; reading starts at offset 0).
    BgzfReader(int fd, int n_threads);
    ~BgzfReader();
    BgzfReader(const BgzfReader&) = delete;
    size_t read(char* dst, size_t n);  // 0 at end of input; throws Error on corrupt data

private:
    struct Block { size_t in_off, in_len, out_off, out_len; };
    bool refill();
    void worker();
    void inflate_block(const Block& b);
    int fd_;
    std::vector<unsigned char> in_;   // compressed bytes: whole blocks, then a partial tail
    size_t in_have_ = 0;
    std::vector<char> out_;
    size_t out_pos_ = 0, out_len_ = 0;
    bool eof_ = false;
    std::vector<Block> blocks_;
    // worker pool
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    uint64_t generation_ = 0;
    std::atomic<size_t> next_{0};
    size_t finished_workers_ = 0;
    bool stop_ = false;
    std::string error_;          // of the lowest-numbered block a worker could not inflate in this round
    size_t first_bad_ = SIZE_MAX;  // its index
    std::string pending_error_;  // raised by the next round: the blocks in front of the broken one go out first
};

}  // namespace mkh
