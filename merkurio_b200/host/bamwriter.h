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
    BamWriter(const std::string& path, const std::vector<std::string>& header_lines);
    ~BamWriter();
    void write_sam_line(const std::string& line);  // one alignment in SAM text form
    void close();

private:
    void put(const void* p, size_t n);
    void flush_block();
    FILE* f_ = nullptr;
    std::vector<uint8_t> buf_;
    std::map<std::string, int32_t> ref_ids_;
};

}  // namespace mkh
