// BAM output for `tag -o out.bam` (the RecordWriter the reference gets from the `bam` crate,
// src/cmd_tag.rs:254-291): SAM text records are encoded to BAM and written as BGZF blocks.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace mkh {

class BamWriter {
public:
    // threads: BGZF blocks are compressed concurrently, 64 at a time (the reference's `-p`, src/cmd_tag.rs:504-507)
    BamWriter(const std::string& path, const std::vector<std::string>& header_lines, int threads = 1);
    ~BamWriter();
    void write_sam_line(const std::string& line);  // one alignment in SAM text form
    // One alignment as it sits in a BAM file (the block_size bytes after the length field) with a Z-typed
    // optional field appended. Only valid if the record's reference ids index this writer's @SQ lines.
    void write_bam_record(const char* body, size_t len, const std::string& tag, const std::string& value);
    const std::vector<std::string>& ref_names() const { return ref_names_; }
    void close();

private:
    void put(const void* p, size_t n);
    void flush_block();
    void compress_pending();
    FILE* f_ = nullptr;
    int threads_ = 1;
    std::vector<uint8_t> buf_;                    // the block being filled
    std::vector<std::vector<uint8_t>> pending_;   // full blocks waiting to be compressed together
    std::map<std::string, int32_t> ref_ids_;
    std::vector<std::string> ref_names_;
};

// Value of an optional field of a BAM record body: 0 = absent, 1 = Z string (value in *val), 2 = other type.
int bam_find_tag(const char* body, size_t len, const std::string& tag, std::string* val);

}  // namespace mkh
