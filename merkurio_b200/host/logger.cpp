#include "logger.h"

#include <charconv>

#include "common.h"

namespace mkh {

std::string json_escape(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            default:
                if (c < 0x20) {
                    char b[8];
                    std::snprintf(b, sizeof b, "\\u%04x", c);
                    o += b;
                } else {
                    o += (char)c;
                }
        }
    }
    return o + "\"";
}

std::string Json::pretty(int indent) const {
    std::string pad(indent + 2, ' '), end(indent, ' ');
    switch (kind) {
        case Null: return "null";
        case Bool: return b ? "true" : "false";
        case Int: return std::to_string(i);
        case Str: return json_escape(s);
        case Arr: {
            if (a.empty()) return "[]";
            std::string r = "[\n";
            for (size_t k = 0; k < a.size(); ++k) r += pad + a[k].pretty(indent + 2) + (k + 1 < a.size() ? ",\n" : "\n");
            return r + end + "]";
        }
        case Obj: {
            if (o.empty()) return "{}";
            std::string r = "{\n";
            size_t k = 0;
            for (auto& kv : o) r += pad + json_escape(kv.first) + ": " + kv.second.pretty(indent + 2) + (++k < o.size() ? ",\n" : "\n");
            return r + end + "}";
        }
    }
    return "null";
}

std::unique_ptr<Sink> Sink::open(const std::string& path, const std::string& what) {
    std::unique_ptr<Sink> s(new Sink);
    if (path == "STDOUT") {
        s->f_ = stdout;
    } else {
        s->f_ = std::fopen(path.c_str(), "wb");
        if (!s->f_) throw Error(what + ": " + path);
        s->owned_ = true;
    }
    return s;
}
Sink::~Sink() {
    if (f_) std::fflush(f_);
    if (f_ && owned_) std::fclose(f_);
}

void BufferedLogger::log_fields(const std::string& prefix, std::string_view record, const std::string& pattern, uint64_t index) {
    if (!sink_) return;  // no text log asked for (the JSON log alone, or none)
    buf_ += prefix; buf_ += '\t';
    buf_ += record; buf_ += '\t';
    buf_ += pattern; buf_ += '\t';
    char digits[24];
    auto r = std::to_chars(digits, digits + sizeof digits, index);
    buf_.append(digits, r.ptr);
    buf_ += '\n';
    if (buf_.size() >= cap_) flush();
}
void BufferedLogger::flush() {
    if (sink_ && !buf_.empty()) sink_->write(buf_);
    buf_.clear();
}

JsonLogger::JsonLogger(std::unique_ptr<Sink> sink, size_t buffer_size) : sink_(std::move(sink)), cap_(buffer_size) {
    if (sink_) sink_->write("{\n  \"matching_records\": [\n");
}
namespace {
// s as a JSON string at the end of out; ordinary text (nothing to escape) is appended in one piece
void append_json_string(std::string& out, std::string_view s) {
    bool plain = true;
    for (unsigned char c : s) plain &= (c >= 0x20 && c != '"' && c != '\\');
    if (plain) {
        out += '"';
        out += s;
        out += '"';
    } else {
        out += json_escape(std::string(s));
    }
}
}  // namespace

void JsonLogger::log_fields(const std::string& file, std::string_view record, const std::string& pattern, uint64_t index) {
    if (!first_) buf_ += ",\n";
    first_ = false;
    // keys in sorted order, 2-space pretty print, every line prefixed with 4 spaces
    buf_ += "    {\n      \"file\": ";
    append_json_string(buf_, file);
    buf_ += ",\n      \"pattern\": ";
    append_json_string(buf_, pattern);
    buf_ += ",\n      \"position\": \"";
    char digits[24];
    auto r = std::to_chars(digits, digits + sizeof digits, index);
    buf_.append(digits, r.ptr);
    buf_ += "\",\n      \"record_id\": ";
    append_json_string(buf_, record);
    buf_ += "\n    }\n";
    if (buf_.size() >= cap_) flush();
}
void JsonLogger::flush() {
    if (sink_ && !buf_.empty()) sink_->write(buf_);
    buf_.clear();
}
void JsonLogger::write_indented_value(const Json& v, int indent) {
    // continuation lines get `indent` extra spaces: the same as pretty-printing at that depth
    buf_ += v.pretty(indent);
    buf_ += '\n';
}
void JsonLogger::finalize(const Json& meta, const std::vector<std::string>& patterns, const std::vector<uint64_t>& counts, const Json& summary,
                          const Json* paired) {
    auto pop_nl = [&] { if (!buf_.empty() && buf_.back() == '\n') buf_.pop_back(); };
    buf_ += "  ],\n  \"meta_information\": ";
    write_indented_value(meta, 2); pop_nl();
    if (paired) {
        buf_ += ",\n  \"paired_end_reads_statistics\": ";
        write_indented_value(*paired, 2); pop_nl();
    }
    // written straight from the list (a million queries would otherwise go through a map and one big string)
    buf_ += ",\n  \"pattern_hit_counts\": ";
    if (patterns.empty()) {
        buf_ += "{}";
    } else {
        buf_ += "{\n";
        for (size_t i = 0; i < patterns.size(); ++i) {
            const std::string& p = patterns[i];
            bool plain = true;  // nothing to escape (every query of ordinary input): no temporary string
            for (unsigned char c : p) plain &= (c >= 0x20 && c != '"' && c != '\\');
            buf_ += "    ";
            if (plain) {
                buf_ += '"';
                buf_ += p;
                buf_ += '"';
            } else {
                buf_ += json_escape(p);
            }
            buf_ += ": ";
            char digits[24];
            auto r = std::to_chars(digits, digits + sizeof digits, counts[i]);
            buf_.append(digits, r.ptr);
            buf_ += i + 1 < patterns.size() ? ",\n" : "\n";
            if (buf_.size() >= (1u << 20)) flush();
        }
        buf_ += "  }";
    }
    buf_ += ",\n  \"summary_statistics\": ";
    write_indented_value(summary, 2); pop_nl();
    buf_ += "\n}\n";
    flush();
    if (sink_) sink_->flush();
}

}  // namespace mkh
