// Host side of the query path, mirroring the reference's src/helpers.rs (names, argument meaning
// and error texts): list parsing, reverse complement / canonical form, algorithm recommendation,
// output naming and the log-flag conflict rules.
#pragma once
#include <optional>
#include <string>
#include <vector>

#include "common.h"

namespace mkh {

// src/helpers.rs:16-25
void error_if_directory(const std::string& path, const std::string& description);
// src/helpers.rs:29-43 — suffix goes before the first dot of the file name
std::string add_suffix_to_file_prefix(const std::string& path, const std::string& suffix);
// src/helpers.rs:48-68
std::string identify_uncompressed_type(const std::string& path);
// src/helpers.rs:139-163
std::vector<std::string> read_kmers_from_file(const std::string& path);
// src/helpers.rs:76-133 — sorted, unique, non-empty; index in the result == pattern id
std::vector<std::string> parse_pattern_list(const std::optional<std::string>& kmer_file,
                                            const std::optional<std::vector<std::string>>& kmer_seq,
                                            bool reverse_complement, bool canonical, bool lowercase, bool uppercase);
// src/helpers.rs:172-200 — "STDOUT" is the value of -l / -j given without an argument
void check_log_flag_conflict(const std::optional<std::string>& out_log, const std::optional<std::string>& json_log,
                             const std::optional<std::string>& out_file, bool suppress_output);
// src/helpers.rs:203-211
bool recommend_aho_corasick(const std::vector<std::string>& patterns);
// src/cmd_extract.rs:165-171 == src/cmd_tag.rs:183-189
bool choose_aho_corasick(const std::vector<std::string>& patterns, bool case_insensitive, const std::optional<size_t>& q_size,
                         bool aho_corasick_flag);
// src/pattern_matching.rs:213-225 and :61-78, src/pattern_preprocessing.rs:31-35: the errors the
// reference raises while building its BNDMq matchers (the device needs neither q nor masks, but the
// failure behaviour of `-q` is part of the interface)
size_t tune_q_value(const std::string& pattern);
void validate_bndmq(const std::vector<std::string>& patterns, const std::optional<size_t>& q_size);

// needletail 0.6.3 (Cargo.lock:382-383): complement of ACGT / IUPAC in both cases, others unchanged
std::string reverse_complement(const std::string& seq);
std::string canonical(const std::string& seq);

std::string timestamp_now();
extern const char* const kProgram;
extern const char* const kVersion;

}  // namespace mkh
