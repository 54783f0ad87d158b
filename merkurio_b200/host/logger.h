// Match-log sinks, byte-compatible with the reference's src/logger.rs: BufferedLogger (TSV text log)
// and JsonLogger (streamed JSON log). They consume the compacted hit lists of the device.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

namespace mkh {

// Minimal JSON value with serde_json's pretty printer (2-space indent, object keys sorted).
struct Json {
    enum Kind { Null, Bool, Int, Str, Arr, Obj } kind = Null;
    bool b = false;
    int64_t i = 0;
    std::string s;
    std::vector<Json> a;
    std::map<std::string, Json> o;
    static Json null() { return Json(); }
    static Json boolean(bool v) { Json j; j.kind = Bool; j.b = v; return j; }
    static Json integer(int64_t v) { Json j; j.kind = Int; j.i = v; return j; }
    static Json string(const std::string& v) { Json j; j.kind = Str; j.s = v; return j; }
    static Json array() { Json j; j.kind = Arr; return j; }
    static Json object() { Json j; j.kind = Obj; return j; }
    Json& operator[](const std::string& k) { kind = Obj; return o[k]; }
    std::string pretty(int indent = 0) const;
};
std::string json_escape(const std::string& s);

// A sink that is a file, stdout, or nothing.
class Sink {
public:
    Sink() = default;
    static std::unique_ptr<Sink> open(const std::string& path_or_STDOUT, const std::string& what);
    ~Sink();
    void write(const std::string& s) { if (f_) std::fwrite(s.data(), 1, s.size(), f_); }
    void write(const char* p, size_t n) { if (f_) std::fwrite(p, 1, n, f_); }
    void flush() { if (f_) std::fflush(f_); }
private:
    FILE* f_ = nullptr;
    bool owned_ = false;
};

// src/logger.rs:11-83 (the reference additionally keeps every line in memory; nothing reads them)
class BufferedLogger {
public:
    BufferedLogger(std::unique_ptr<Sink> sink, size_t buffer_size) : sink_(std::move(sink)), cap_(buffer_size) { buf_.reserve(buffer_size + 256); }
    bool active() const { return (bool)sink_; }
    void log_fields(const std::string& prefix, std::string_view record, const std::string& pattern, uint64_t index);
    void write_header(const std::string& header) { if (sink_) sink_->write(header); }
    void flush();
private:
    std::unique_ptr<Sink> sink_;
    std::string buf_;
    size_t cap_;
};

// src/logger.rs:86-191
class JsonLogger {
public:
    JsonLogger(std::unique_ptr<Sink> sink, size_t buffer_size);
    void log_fields(const std::string& file, std::string_view record, const std::string& pattern, uint64_t index);
    void flush();
    // patterns: the sorted unique query list (= the key order serde_json's sorted map gives), counts[i] its hits
    void finalize(const Json& meta_information, const std::vector<std::string>& patterns, const std::vector<uint64_t>& counts,
                  const Json& summary_statistics, const Json* paired_end_stats);
private:
    void write_indented_value(const Json& v, int indent);
    std::unique_ptr<Sink> sink_;
    std::string buf_;
    size_t cap_;
    bool first_ = true;
};

}  // namespace mkh
