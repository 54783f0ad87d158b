// Shared host utilities: anyhow-style errors, Rust-compatible path helpers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace mkh {

// An error bubbling to main. The reference prints `Error: <outermost>` followed by a
// `Caused by:` chain (anyhow's Debug output) and exits with status 1.
class Error : public std::exception {
public:
    explicit Error(std::string msg) { chain_.push_back(std::move(msg)); }
    Error with_context(std::string ctx) const {
        Error e(*this);
        e.chain_.insert(e.chain_.begin(), std::move(ctx));
        return e;
    }
    const std::vector<std::string>& chain() const { return chain_; }
    const char* what() const noexcept override { return chain_.front().c_str(); }
    std::string report() const {
        std::string s = "Error: " + chain_.front() + "\n";
        if (chain_.size() > 1) {
            s += "\nCaused by:\n";
            for (size_t i = 1; i < chain_.size(); ++i)
                s += (chain_.size() > 2 ? "    " + std::to_string(i - 1) + ": " : "    ") + chain_[i] + "\n";
        }
        return s;
    }

private:
    std::vector<std::string> chain_;
};

// --- record output: every write is checked -------------------------------------------------------
// A full disk, a quota or a closed pipe must end the run with an error and a non-zero status instead of a truncated
// output file: the reference propagates record-write failures ("Error writing record to output file",
// src/cmd_tag.rs:494-496; extract unwraps the result of record.write, src/cmd_extract.rs:403,603-604). Only its
// loggers ignore write errors (src/logger.rs), and so do ours.
void write_all(FILE* f, const void* p, size_t n);  // fwrite, throws Error on a short write
void flush_checked(FILE* f);                       // fflush, throws Error
void close_checked(FILE* f);                       // fclose, throws Error

// --- std::path::Path semantics the reference relies on -----------------------------------------
std::string path_file_name(const std::string& path);                        // Path::file_name
bool path_extension(const std::string& path, std::string* ext);             // Path::extension
std::string path_with_extension(const std::string& path, const std::string& ext);  // with_extension
bool path_is_dir(const std::string& path);
bool path_exists(const std::string& path);
std::string rust_debug_string(const std::string& s);                        // {:?} of a str / Path

}  // namespace mkh
