// The two subcommands, with the reference's argument structs (src/cmd_extract.rs:64-141,
// src/cmd_tag.rs:65-150): same field names and meaning. "STDOUT" is the value -l / -j take when given
// without an argument.
#pragma once
#include <optional>
#include <string>
#include <vector>

namespace mkh {

struct CmdExtract {
    std::string in_fastx;
    std::optional<std::string> in_fastq_2;
    std::optional<std::vector<std::string>> kmer_seq;
    std::optional<std::string> kmer_file;
    std::optional<std::string> out_fastx;
    bool reverse_complement = false, canonical = false;
    std::optional<std::string> out_log, json_log;
    bool suppress_output = false, invert_match = false, case_insensitive = false, lowercase = false, uppercase = false;
    std::optional<size_t> q_size;
    bool aho_corasick = false;
    std::vector<std::string> argv;  // env::args(), for the logs
};

struct CmdTag {
    std::string in_file;
    std::optional<std::string> out_file;
    std::optional<std::vector<std::string>> kmer_seq;
    std::optional<std::string> kmer_file;
    bool reverse_complement = false, canonical = false;
    std::string tag = "km";
    std::optional<std::string> out_log, json_log;
    int threads = 1;
    bool suppress_output = false, filter_matching = false, invert_match = false, case_insensitive = false, lowercase = false,
         uppercase = false;
    std::optional<size_t> q_size;
    bool aho_corasick = false;
    std::vector<std::string> argv;
};

void extract_records(CmdExtract args);  // src/cmd_extract.rs:143
void tag_records(CmdTag args);          // src/cmd_tag.rs:155

}  // namespace mkh
