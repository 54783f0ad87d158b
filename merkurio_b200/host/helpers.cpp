#include "helpers.h"

#include <cerrno>
#include <cstring>

#include <sys/stat.h>

#include <algorithm>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <thread>

namespace mkh {

const char* const kProgram = "merkurio";
const char* const kVersion = "1.0.0";  // crate_version!() of the reference tree (Cargo.toml:3)

// ---------------------------------------------------------------------------------------------
// paths
// ---------------------------------------------------------------------------------------------
static Error write_error() {
    const int e = errno;
    return Error(std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")").with_context("Error writing record to output file");
}
void write_all(FILE* f, const void* p, size_t n) {
    if (n && std::fwrite(p, 1, n, f) != n) throw write_error();
}
void flush_checked(FILE* f) {
    if (std::fflush(f) != 0) throw write_error();
}
void close_checked(FILE* f) {
    if (std::fclose(f) != 0) throw write_error();
}

std::string path_file_name(const std::string& path) {
    std::string p = path;
    while (p.size() > 1 && p.back() == '/') p.pop_back();
    size_t i = p.rfind('/');
    return i == std::string::npos ? p : p.substr(i + 1);
}

bool path_extension(const std::string& path, std::string* ext) {
    std::string name = path_file_name(path);
    if (name.empty() || name == "..") return false;
    size_t i = name.rfind('.');
    if (i == std::string::npos || i == 0) return false;
    *ext = name.substr(i + 1);
    return true;
}

std::string path_with_extension(const std::string& path, const std::string& ext) {
    std::string name = path_file_name(path);
    std::string dir = path.substr(0, path.size() - name.size());
    size_t i = name.rfind('.');
    std::string stem = (i == std::string::npos || i == 0) ? name : name.substr(0, i);
    return dir + stem + (ext.empty() ? "" : "." + ext);
}

bool path_is_dir(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool path_exists(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0;
}

std::string rust_debug_string(const std::string& s) {
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += c; }
        else if (c == '\n') o += "\\n";
        else if (c == '\t') o += "\\t";
        else if (c == '\r') o += "\\r";
        else o += c;
    }
    return o + "\"";
}

// ---------------------------------------------------------------------------------------------
void error_if_directory(const std::string& path, const std::string& description) {
    if (path_is_dir(path)) throw Error(description + " '" + path + "' is a directory, not a file.");
}

std::string add_suffix_to_file_prefix(const std::string& path, const std::string& suffix) {
    std::string name = path_file_name(path);
    std::string dir = path.substr(0, path.size() - name.size());
    size_t i = name.find('.');
    if (i == std::string::npos) return dir + name + suffix;
    return dir + name.substr(0, i) + suffix + name.substr(i);
}

std::string identify_uncompressed_type(const std::string& path) {
    if (path_is_dir(path)) throw Error("The path points to a directory.");
    std::string ext;
    if (!path_extension(path, &ext)) throw Error("Path has no extension");
    if (ext == "gz" || ext == "bz" || ext == "bz2" || ext == "xz") {
        std::string inner;
        if (!path_extension(path_with_extension(path, ""), &inner)) throw Error("Could not determine uncompressed file type");
        return inner;
    }
    return ext;
}

static bool is_rust_whitespace(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); }

std::vector<std::string> read_kmers_from_file(const std::string& path) {
    if (path_is_dir(path)) throw Error("K-mer file path '" + path + "' is a directory, not a file.");
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        Error inner(path_exists(path) ? "Permission denied (os error 13)" : "No such file or directory (os error 2)");
        throw inner.with_context(path_exists(path) ? "Error reading file: " + path : "File not found.");
    }
    in.seekg(0, std::ios::end);
    const std::streamoff size = in.tellg();
    std::string content;
    if (size > 0) {
        content.resize((size_t)size);
        in.seekg(0, std::ios::beg);
        in.read(&content[0], size);
        content.resize((size_t)std::max<std::streamsize>(in.gcount(), 0));
    } else {  // not seekable (a pipe): read it as a stream
        in.clear();
        std::stringstream ss;
        ss << in.rdbuf();
        content = ss.str();
    }
    std::vector<std::string> kmers;
    kmers.reserve((size_t)std::count(content.begin(), content.end(), '\n') + 1);
    size_t pos = 0;
    while (pos < content.size()) {  // str::lines(): split on '\n', strip one trailing '\r'
        const char* nlp = static_cast<const char*>(std::memchr(content.data() + pos, '\n', content.size() - pos));
        size_t end = nlp ? (size_t)(nlp - content.data()) : content.size();
        const size_t next = nlp ? end + 1 : content.size();
        if (end > pos && content[end - 1] == '\r') --end;
        if (end > pos && content[pos] != '#' && content[pos] != '>') {
            size_t a = pos, b = end;  // trim() in place: one allocation per line
            while (a < b && is_rust_whitespace((unsigned char)content[a])) ++a;
            while (b > a && is_rust_whitespace((unsigned char)content[b - 1])) --b;
            kmers.emplace_back(content.data() + a, b - a);
        }
        pos = next;
    }
    if (kmers.empty()) throw Error("No k-mers found in the file.");
    return kmers;
}

static unsigned char complement(unsigned char c) {
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
        case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
        case 'r': return 'y'; case 'y': return 'r'; case 'k': return 'm'; case 'm': return 'k';
        case 'b': return 'v'; case 'v': return 'b'; case 'd': return 'h'; case 'h': return 'd';
        default: return c;  // S, W, N and every other byte are their own complement
    }
}

std::string reverse_complement(const std::string& seq) {
    std::string out(seq.size(), '\0');
    for (size_t i = 0; i < seq.size(); ++i) out[seq.size() - 1 - i] = (char)complement((unsigned char)seq[i]);
    return out;
}

std::string canonical(const std::string& seq) {
    std::string rc = reverse_complement(seq);
    return rc < seq ? rc : seq;  // ties keep the original
}

static std::string ascii_case(const std::string& s, bool lower) {
    std::string o = s;
    for (char& c : o) {
        if (lower && c >= 'A' && c <= 'Z') c = (char)(c | 0x20);
        if (!lower && c >= 'a' && c <= 'z') c = (char)(c & ~0x20);
    }
    return o;
}

std::vector<std::string> parse_pattern_list(const std::optional<std::string>& kmer_file,
                                            const std::optional<std::vector<std::string>>& kmer_seq,
                                            bool reverse_complement_, bool canonical_, bool lowercase, bool uppercase) {
    std::vector<std::string> pats;
    if (kmer_file) {
        try {
            pats = read_kmers_from_file(*kmer_file);
        } catch (const Error& e) {
            throw e.with_context("Problem reading k-mers from file: " + rust_debug_string(*kmer_file));
        }
    } else {
        if (!kmer_seq) throw Error("No k-mer sequence provided.");
        pats = *kmer_seq;
    }
    if (lowercase) for (auto& p : pats) p = ascii_case(p, true);
    else if (uppercase) for (auto& p : pats) p = ascii_case(p, false);
    if (reverse_complement_) {
        size_t n = pats.size();
        for (size_t i = 0; i < n; ++i) pats.push_back(reverse_complement(pats[i]));
    }
    if (canonical_) for (auto& p : pats) p = canonical(p);
    pats.erase(std::remove_if(pats.begin(), pats.end(), [](const std::string& s) { return s.empty(); }), pats.end());
    auto bytewise = [](const std::string& a, const std::string& b) {
        int c = std::memcmp(a.data(), b.data(), std::min(a.size(), b.size()));
        return c != 0 ? c < 0 : a.size() < b.size();
    };
    // large lists (cfg5: a million queries): runs sorted on several threads, then merged pairwise
    const size_t n_threads = pats.size() < 100000 ? 1 : std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 8);
    if (n_threads == 1) {
        std::sort(pats.begin(), pats.end(), bytewise);
    } else {
        std::vector<size_t> cut(n_threads + 1);
        for (size_t i = 0; i <= n_threads; ++i) cut[i] = pats.size() * i / n_threads;
        {
            std::vector<std::thread> th;
            for (size_t i = 0; i < n_threads; ++i)
                th.emplace_back([&, i] { std::sort(pats.begin() + (std::ptrdiff_t)cut[i], pats.begin() + (std::ptrdiff_t)cut[i + 1], bytewise); });
            for (auto& t : th) t.join();
        }
        for (size_t width = 1; width < n_threads; width *= 2) {
            std::vector<std::thread> th;
            for (size_t i = 0; i + width < n_threads; i += 2 * width)
                th.emplace_back([&, i, width] {
                    std::inplace_merge(pats.begin() + (std::ptrdiff_t)cut[i], pats.begin() + (std::ptrdiff_t)cut[i + width],
                                       pats.begin() + (std::ptrdiff_t)cut[std::min(i + 2 * width, n_threads)], bytewise);
                });
            for (auto& t : th) t.join();
        }
    }
    pats.erase(std::unique(pats.begin(), pats.end()), pats.end());
    if (pats.empty()) throw Error("No k-mers found in file or provided sequence.");
    return pats;
}

void check_log_flag_conflict(const std::optional<std::string>& out_log, const std::optional<std::string>& json_log,
                             const std::optional<std::string>& out_file, bool suppress_output) {
    bool l = out_log && *out_log == "STDOUT", j = json_log && *json_log == "STDOUT";
    if (l && j)
        throw Error("Cannot use both -l/--out-log and -j/--json-log with no arguments (both to stdout). Please specify a file for at least one.");
    if ((l || j) && !out_file && !suppress_output)
        throw Error("Cannot write log to stdout when normal output is also stdout. Specify an output file with -o or suppress output with -S.");
}

bool recommend_aho_corasick(const std::vector<std::string>& patterns) {
    size_t max_len = 0;
    for (auto& p : patterns) max_len = std::max(max_len, p.size());
    return patterns.size() >= 14 || max_len > 64;
}

bool choose_aho_corasick(const std::vector<std::string>& patterns, bool case_insensitive, const std::optional<size_t>& q_size,
                         bool aho_corasick_flag) {
    if (case_insensitive) return true;
    if (!q_size && !aho_corasick_flag) return recommend_aho_corasick(patterns);
    return aho_corasick_flag;
}

size_t tune_q_value(const std::string& pattern) {
    size_t n = pattern.size();
    if (n <= 1) return 1;
    if (n <= 3) return 2;
    if (n <= 8) return 3;
    if (n <= 30) return 4;
    if (n <= 55) return 5;
    if (n <= 64) return 6;
    throw Error("Pattern length is too long for BNDMq.");
}

void validate_bndmq(const std::vector<std::string>& patterns, const std::optional<size_t>& q_size) {
    for (auto& p : patterns) {
        size_t q = q_size ? *q_size : tune_q_value(p);
        if (p.empty()) throw Error("Pattern is empty.");
        if (q == 0 || q > p.size()) throw Error("Invalid q-gram length: " + std::to_string(q) + ". Must be between 1 and pattern length.");
        if (p.size() > 64)
            throw Error("Pattern length " + std::to_string(p.size()) + " is too large for this architecture when using BNDM (max 64).");
    }
}

std::string timestamp_now() {
    std::time_t t = std::time(nullptr);
    std::tm tm{};
    localtime_r(&t, &tm);
    char buf[64], off[16];
    std::strftime(buf, sizeof buf, "%Y-%m-%dT%H:%M:%S", &tm);
    std::strftime(off, sizeof off, "%z", &tm);  // +hhmm
    std::string o = off;
    if (o.size() == 5) o = o.substr(0, 3) + ":" + o.substr(3);
    const char* tz = tm.tm_zone ? tm.tm_zone : "UTC";
    return std::string(buf) + o + "[" + tz + "]";
}

}  // namespace mkh
