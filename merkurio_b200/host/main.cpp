// merkurio — command line with the reference's grammar (src/main.rs:14-54 and the clap structs in
// src/cmd_extract.rs:33-141, src/cmd_tag.rs:29-150): the same subcommands, flags, value counts,
// argument groups and exit codes (2 for usage errors, 1 for run-time errors, as clap / anyhow do).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <unistd.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "commands.h"
#include "device.h"
#include "aln_stream.h"
#include "fastq_stream.h"
#include "helpers.h"
#include "io.h"

using namespace mkh;

namespace {

enum Kind { FLAG, VALUE, OPT_VALUE, MULTI };
struct Opt {
    char short_name;
    const char* long_name;
    Kind kind;
    const char* value_name;
    const char* help;
};

const Opt kExtract[] = {
    {'i', "in-fastx", VALUE, "IN_FASTX", "Input path for (compressed) FASTQ/A file"},
    {'2', "in-fastq-2", VALUE, "IN_FASTQ_2", "Input path for second FASTQ file (only for paired-end read processing)"},
    {'s', "kmer-seq", MULTI, "KMER_SEQ", "Query sequences (accepts multiple sequences after the flag, separated by a space)"},
    {'f', "kmer-file", VALUE, "KMER_FILE", "Input path for file containing list of k-mers, one per line"},
    {'o', "out-fastx", VALUE, "OUT_FASTX", "Output file path for FASTQ/A file (extension derived from input file)"},
    {'r', "reverse-complement", FLAG, nullptr, "Also search for reverse complements of k-mers"},
    {'c', "canonical", FLAG, nullptr, "Search only for the canonical forms of k-mers"},
    {'l', "out-log", OPT_VALUE, "OUT_LOG", "Print detailed match information to stdout, or to a file if a path is provided"},
    {'j', "json-log", OPT_VALUE, "JSON_LOG", "Write JSON log to stdout, or to a file if a path is provided"},
    {'S', "suppress-output", FLAG, nullptr, "Suppress output of found records"},
    {'v', "invert-match", FLAG, nullptr, "Invert the sense of matching, to select non-matching records"},
    {'I', "case-insensitive", FLAG, nullptr, "Use case-insensitive matching"},
    {'L', "lowercase", FLAG, nullptr, "Convert all input sequences to lowercase"},
    {'U', "uppercase", FLAG, nullptr, "Convert all input sequences to uppercase"},
    {'q', "q-size", VALUE, "Q_SIZE", "Manually set size of q-grams (BNDMq report order and count semantics)"},
    {'a', "aho-corasick", FLAG, nullptr, "Aho-Corasick report order and count semantics"},
};
const Opt kTag[] = {
    {'i', "in-file", VALUE, "IN_FILE", "Input path for SAM/BAM file"},
    {'o', "out-file", VALUE, "OUT_FILE", "Output path for SAM/BAM file with annotations"},
    {'s', "kmer-seq", MULTI, "KMER_SEQ", "Query sequences (accepts multiple sequences after the flag, separated by a space)"},
    {'f', "kmer-file", VALUE, "KMER_FILE", "Input path for file containing list of k-mers, one per line"},
    {'r', "reverse-complement", FLAG, nullptr, "Also search for reverse complements of k-mers"},
    {'c', "canonical", FLAG, nullptr, "Search only for the canonical forms of k-mers"},
    {'t', "tag", VALUE, "TAG", "Tag to add to the SAM/BAM file with the presence of k-mers [default: km]"},
    {'l', "out-log", OPT_VALUE, "OUT_LOG", "Print detailed match information to stdout, or to a file if a path is provided"},
    {'j', "json-log", OPT_VALUE, "JSON_LOG", "Write JSON log to stdout, or to a file if a path is provided"},
    {'p', "threads", VALUE, "THREADS", "Number of parallel threads to use for processing BAM files [default: 1]"},
    {'S', "suppress-output", FLAG, nullptr, "Suppress output of found records"},
    {'m', "filter-matching", FLAG, nullptr, "Filter records to keep only those with matching k-mers"},
    {'v', "invert-match", FLAG, nullptr, "Invert the sense of matching"},
    {'I', "case-insensitive", FLAG, nullptr, "Use case-insensitive matching"},
    {'L', "lowercase", FLAG, nullptr, "Convert all input sequences to lowercase"},
    {'U', "uppercase", FLAG, nullptr, "Convert all input sequences to uppercase"},
    {'q', "q-size", VALUE, "Q_SIZE", "Manually set size of q-grams (BNDMq report order and count semantics)"},
    {'a', "aho-corasick", FLAG, nullptr, "Aho-Corasick report order and count semantics"},
};

[[noreturn]] void usage_error(const std::string& msg, const char* sub) {
    std::fprintf(stderr, "error: %s\n\nUsage: merkurio %s [OPTIONS]\n\nFor more information, try '--help'.\n", msg.c_str(), sub);
    std::exit(2);
}

void print_help(const char* sub, const Opt* opts, size_t n) {
    std::printf("Usage: merkurio %s [OPTIONS]\n\nOptions:\n", sub);
    for (size_t i = 0; i < n; ++i) {
        std::string left = std::string("  -") + opts[i].short_name + ", --" + opts[i].long_name;
        if (opts[i].kind == VALUE) left += std::string(" <") + opts[i].value_name + ">";
        if (opts[i].kind == OPT_VALUE) left += std::string(" [<") + opts[i].value_name + ">]";
        if (opts[i].kind == MULTI) left += std::string(" <") + opts[i].value_name + ">...";
        std::printf("%-40s %s\n", left.c_str(), opts[i].help);
    }
    std::printf("  -h, --help                               Print help\n");
}

std::string display(const Opt& o) {
    std::string s = std::string("--") + o.long_name;
    if (o.kind == VALUE) s += std::string(" <") + o.value_name + ">";
    if (o.kind == OPT_VALUE) s += std::string(" [<") + o.value_name + ">]";
    if (o.kind == MULTI) s += std::string(" <") + o.value_name + ">...";
    return s;
}

struct Parsed {
    std::map<std::string, std::vector<std::string>> values;  // long name -> values (flags: empty vector)
    bool has(const char* k) const { return values.count(k) > 0; }
    const std::string& one(const char* k) const { return values.at(k).front(); }
};

Parsed parse(int argc, char** argv, int first, const char* sub, const Opt* opts, size_t n) {
    Parsed p;
    auto find_long = [&](const std::string& name) -> const Opt* {
        for (size_t i = 0; i < n; ++i) if (name == opts[i].long_name) return &opts[i];
        return nullptr;
    };
    auto find_short = [&](char c) -> const Opt* {
        if (std::strcmp(sub, "extract") == 0 && c == '1') c = 'i';  // short_alias = '1'
        for (size_t i = 0; i < n; ++i) if (c == opts[i].short_name) return &opts[i];
        return nullptr;
    };
    int i = first;
    auto looks_like_flag = [](const char* s) { return s[0] == '-' && s[1] != '\0'; };
    auto take = [&](const Opt& o, const std::string* attached) {
        if (p.has(o.long_name) && o.kind != MULTI) usage_error("the argument '" + display(o) + "' cannot be used multiple times", sub);
        std::vector<std::string>& v = p.values[o.long_name];
        switch (o.kind) {
            case FLAG:
                if (attached) usage_error(std::string("unexpected value '") + *attached + "' for '--" + o.long_name + "' found; no more were expected", sub);
                break;
            case VALUE:
                if (attached) v.push_back(*attached);
                else if (i < argc && !(looks_like_flag(argv[i]) && !(argv[i][1] >= '0' && argv[i][1] <= '9'))) v.push_back(argv[i++]);
                else usage_error("a value is required for '" + display(o) + "' but none was supplied", sub);
                break;
            case OPT_VALUE:
                if (attached) v.push_back(*attached);
                else if (i < argc && !looks_like_flag(argv[i])) v.push_back(argv[i++]);
                else v.push_back("STDOUT");
                break;
            case MULTI:
                if (attached) v.push_back(*attached);
                while (i < argc && !looks_like_flag(argv[i])) v.push_back(argv[i++]);
                if (v.empty()) usage_error("a value is required for '" + display(o) + "' but none was supplied", sub);
                break;
        }
    };
    while (i < argc) {
        std::string tok = argv[i++];
        if (tok == "-h" || tok == "--help") { print_help(sub, opts, n); std::exit(0); }
        if (tok.rfind("--", 0) == 0) {
            std::string name = tok.substr(2), val;
            bool has_val = false;
            size_t eq = name.find('=');
            if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
            const Opt* o = find_long(name);
            if (!o) usage_error("unexpected argument '--" + name + "' found", sub);
            take(*o, has_val ? &val : nullptr);
        } else if (tok.size() >= 2 && tok[0] == '-') {
            for (size_t k = 1; k < tok.size(); ++k) {
                const Opt* o = find_short(tok[k]);
                if (!o) usage_error(std::string("unexpected argument '-") + tok[k] + "' found", sub);
                if (o->kind != FLAG && k + 1 < tok.size()) {
                    std::string rest = tok.substr(k + 1);
                    if (!rest.empty() && rest[0] == '=') rest = rest.substr(1);
                    take(*o, &rest);
                    break;
                }
                take(*o, nullptr);
            }
        } else {
            usage_error("unexpected argument '" + tok + "' found", sub);
        }
    }
    return p;
}

void exclusive(const Parsed& p, const Opt* opts, size_t n, std::initializer_list<const char*> names, const char* sub) {
    const Opt* seen = nullptr;
    for (const char* nm : names) {
        if (!p.has(nm)) continue;
        const Opt* o = nullptr;
        for (size_t i = 0; i < n; ++i) if (std::strcmp(opts[i].long_name, nm) == 0) o = &opts[i];
        if (seen) usage_error("the argument '" + display(*seen) + "' cannot be used with '" + display(*o) + "'", sub);
        seen = o;
    }
}

size_t parse_usize(const std::string& s, const char* what, const char* sub) {
    char* end = nullptr;
    if (s.empty() || s[0] == '-') usage_error("invalid value '" + s + "' for '" + what + "': invalid digit found in string", sub);
    unsigned long long v = std::strtoull(s.c_str(), &end, 10);
    if (*end) usage_error("invalid value '" + s + "' for '" + what + "': invalid digit found in string", sub);
    return (size_t)v;
}

std::optional<std::string> opt_str(const Parsed& p, const char* k) {
    if (!p.has(k)) return std::nullopt;
    return p.one(k);
}

int run(int argc, char** argv) {
    const char* top_help =
        "SeqKatcher has two subcommands, 'extract' and 'tag'. This build runs the matching on NVIDIA B200 GPUs.\n\n"
        "Usage: merkurio <COMMAND>\n\nCommands:\n"
        "  extract  Search for query sequences in FASTA/Q files and extract records containing the patterns\n"
        "  tag      Tag records in a BAM/SAM file with the presence of query sequences\n"
        "  help     Print this message\n\nOptions:\n  -h, --help     Print help\n  -V, --version  Print version\n";
    if (argc < 2) { std::fputs(top_help, stderr); return 2; }
    std::string cmd = argv[1];
    if (cmd == "-h" || cmd == "--help" || cmd == "help") { std::fputs(top_help, stdout); return 0; }
    if (cmd == "-V" || cmd == "--version") { std::printf("%s %s\n", kProgram, kVersion); return 0; }
    std::vector<std::string> all(argv, argv + argc);
    if (cmd == "extract") {
        const size_t n = sizeof kExtract / sizeof kExtract[0];
        if (argc == 2) { print_help("extract", kExtract, n); return 2; }
        Parsed p = parse(argc, argv, 2, "extract", kExtract, n);
        exclusive(p, kExtract, n, {"kmer-seq", "kmer-file"}, "extract");
        exclusive(p, kExtract, n, {"q-size", "aho-corasick"}, "extract");
        exclusive(p, kExtract, n, {"case-insensitive", "lowercase", "uppercase"}, "extract");
        exclusive(p, kExtract, n, {"canonical", "reverse-complement"}, "extract");
        exclusive(p, kExtract, n, {"suppress-output", "out-fastx"}, "extract");
        std::string missing;
        if (!p.has("in-fastx")) missing += "\n  --in-fastx <IN_FASTX>";
        if (!p.has("kmer-seq") && !p.has("kmer-file")) missing += "\n  <--kmer-seq <KMER_SEQ>...|--kmer-file <KMER_FILE>>";
        if (p.has("suppress-output") && !p.has("out-log") && !p.has("json-log")) missing += "\n  <--out-log [<OUT_LOG>]|--json-log [<JSON_LOG>]>";
        if (!missing.empty()) usage_error("the following required arguments were not provided:" + missing, "extract");
        CmdExtract a;
        a.in_fastx = p.one("in-fastx");
        a.in_fastq_2 = opt_str(p, "in-fastq-2");
        if (p.has("kmer-seq")) a.kmer_seq = p.values.at("kmer-seq");
        a.kmer_file = opt_str(p, "kmer-file");
        a.out_fastx = opt_str(p, "out-fastx");
        a.reverse_complement = p.has("reverse-complement");
        a.canonical = p.has("canonical");
        a.out_log = opt_str(p, "out-log");
        a.json_log = opt_str(p, "json-log");
        a.suppress_output = p.has("suppress-output");
        a.invert_match = p.has("invert-match");
        a.case_insensitive = p.has("case-insensitive");
        a.lowercase = p.has("lowercase");
        a.uppercase = p.has("uppercase");
        if (p.has("q-size")) a.q_size = parse_usize(p.one("q-size"), "--q-size <Q_SIZE>", "extract");
        a.aho_corasick = p.has("aho-corasick");
        a.argv = all;
        extract_records(std::move(a));
        return 0;
    }
    if (cmd == "patterns") {
        // diagnostic (not in the reference): print the query list after preprocessing and the algorithm
        // the reference would choose — the host half of the hot path, testable without a GPU
        const size_t n = sizeof kExtract / sizeof kExtract[0];
        Parsed p = parse(argc, argv, 2, "patterns", kExtract, n);
        exclusive(p, kExtract, n, {"kmer-seq", "kmer-file"}, "patterns");
        exclusive(p, kExtract, n, {"canonical", "reverse-complement"}, "patterns");
        exclusive(p, kExtract, n, {"case-insensitive", "lowercase", "uppercase"}, "patterns");
        std::optional<std::vector<std::string>> seqs;
        if (p.has("kmer-seq")) seqs = p.values.at("kmer-seq");
        std::optional<size_t> q;
        if (p.has("q-size")) q = parse_usize(p.one("q-size"), "--q-size <Q_SIZE>", "patterns");
        std::vector<std::string> pats;
        try {
            pats = parse_pattern_list(opt_str(p, "kmer-file"), seqs, p.has("reverse-complement"), p.has("canonical"), p.has("lowercase"), p.has("uppercase"));
        } catch (const Error& e) {
            throw e.with_context("Problem parsing pattern list.");
        }
        bool ac = choose_aho_corasick(pats, p.has("case-insensitive"), q, p.has("aho-corasick"));
        if (!ac) validate_bndmq(pats, q);
        for (auto& s : pats) std::printf("%s\n", s.c_str());
        std::printf("#search_algorithm\t%s\n", ac ? "Aho-Corasick" : "BNDMq");
        return 0;
    }
    if (cmd == "records") {
        // diagnostic (not in the reference): merkurio records <file> <generic|chunked|count|count-generic|cat> [chunk_bytes]
        // dumps what the record readers hand to the matcher and to the writer — the ingest half of the
        // extract path, testable without a GPU. One block per record: id, sequence, then the bytes the
        // FASTA/FASTQ writer would emit; a parse error ends the dump with "#error".
        if (argc < 4) { std::fputs("usage: merkurio records <file> <generic|chunked> [chunk_bytes]\n", stderr); return 2; }
        const std::string path = argv[2], how = argv[3];
        std::string out;
        auto dump = [&](const std::string& id, const std::string& seq, const FastxRecord& r) {
            out += "#id\t" + id + "\n#seq\t" + seq + "\n";
            r.write(&out);
        };
        if (how == "cat") {  // the decompressed byte stream (codecs.cpp), as the readers see it
            std::vector<char> buf(argc > 4 ? (size_t)std::strtoull(argv[4], nullptr, 10) : (size_t)8 << 20);
            try {
                std::unique_ptr<InputStream> in = InputStream::open(path);
                while (size_t got = in->read(buf.data(), buf.size())) std::fwrite(buf.data(), 1, got, stdout);
            } catch (const Error& e) {
                std::fflush(stdout);
                std::fprintf(stderr, "#error\t%s\n", e.what());
                return 1;
            }
            return 0;
        }
        if (how == "count" || how == "count-generic") {  // reader throughput: records and bases only
            uint64_t n = 0, bases = 0;
            if (how == "count") {
                FastqChunkReader cr(path, argc > 4 ? (size_t)std::strtoull(argv[4], nullptr, 10) : (size_t)16 << 20);
                while (std::shared_ptr<Chunk> ch = cr.next())
                    for (const RecSpan& sp : ch->recs) { ++n; bases += sp.seq_len; }
            } else {
                FastxReader rd(path);
                FastxRecord r;
                while (rd.next(&r)) { ++n; bases += r.seq.size(); }
            }
            std::printf("%llu\t%llu\n", (unsigned long long)n, (unsigned long long)bases);
            return 0;
        }
        if (how == "generic") {
            FastxReader rd(path);
            FastxRecord r;
            try {
                while (rd.next(&r)) dump(r.id, r.seq, r);
            } catch (const Error& e) {
                out += std::string("#error\t") + e.what() + "\n";
            }
        } else {
            if (!looks_like_fastq(path)) { std::fputs("#not-fastq\n", stdout); return 0; }
            FastqChunkReader cr(path, argc > 4 ? (size_t)std::strtoull(argv[4], nullptr, 10) : (size_t)16 << 20);
            try {
            while (std::shared_ptr<Chunk> ch = cr.next()) {
                for (const RecSpan& sp : ch->recs) {
                    FastxRecord r;
                    r.id.assign(ch->id(sp), sp.id_len);
                    r.seq.assign(ch->seq(sp), sp.seq_len);
                    r.raw = r.seq;
                    r.qual.assign(ch->qual(sp), sp.seq_len);
                    r.fastq = true;
                    r.crlf = sp.crlf;
                    if (sp.plain) {
                        out += "#id\t" + r.id + "\n#seq\t" + r.seq + "\n";
                        out.append(ch->data.data() + sp.start, sp.end - sp.start);
                    } else {
                        dump(r.id, r.seq, r);
                    }
                }
                if (ch->failed) out += "#error\tError during FASTQ/A record parsing.\n";
            }
            } catch (const Error& e) {  // read / decompression error: reported like the line reader's
                out += std::string("#error\t") + e.what() + "\n";
            }
        }
        std::fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    }
    if (cmd == "alnrecords") {
        // diagnostic (not in the reference): merkurio alnrecords <file.sam|file.bam> <generic|chunked> [chunk_bytes]
        // dumps what the alignment readers hand to the matcher (name, length, 4-bit sequence) and to the
        // writer (the SAM text of the record) — the ingest half of the tag path, testable without a GPU.
        if (argc < 4) { std::fputs("usage: merkurio alnrecords <file> <generic|chunked> [chunk_bytes]\n", stderr); return 2; }
        const std::string path = argv[2], how = argv[3];
        const bool bam = path.size() > 4 && path.compare(path.size() - 4, 4, ".bam") == 0;
        std::string out;
        auto dump = [&](const std::string& name, uint32_t l_seq, const uint8_t* packed, const std::string& line) {
            out += "#name\t" + name + "\t" + std::to_string(l_seq) + "\t";
            for (uint32_t i = 0; i < l_seq; ++i) out += kNibbleChars[(packed[i >> 1] >> ((i & 1) ? 0 : 4)) & 0xF];
            out += "\n" + line + "\n";
        };
        std::unique_ptr<AlnReader> rd(new AlnReader(path, bam));
        for (auto& h : rd->header_lines()) out += "#header\t" + h + "\n";
        if (how == "generic") {
            AlnRecord r;
            try {
                while (rd->next(&r)) dump(r.name, r.l_seq, r.packed.data(), r.sam_line);
            } catch (const Error& e) {
                out += std::string("#error\t") + e.what() + "\n";
            }
        } else {
            const std::vector<std::string> refs = rd->refs();
            AlnChunkReader cr(std::move(rd), argc > 4 ? (size_t)std::strtoull(argv[4], nullptr, 10) : (size_t)8 << 20);
            while (std::shared_ptr<AlnChunk> ch = cr.next()) {
                for (const AlnSpan& sp : ch->recs) {
                    std::string name(ch->data.data() + sp.name_off, sp.name_len), line;
                    std::vector<uint8_t> packed((sp.l_seq + 1) / 2, 0);
                    if (ch->bam) {
                        std::string nm;
                        bam_body_to_sam(ch->data.data() + sp.off, sp.len, refs, &nm, &line, nullptr, nullptr);
                        if (!packed.empty()) std::memcpy(packed.data(), ch->data.data() + sp.seq_off, packed.size());
                    } else {
                        line.assign(ch->data.data() + sp.off, sp.len);
                        for (uint32_t i = 0; i < sp.l_seq; ++i)
                            packed[i >> 1] |= (uint8_t)(nibble_of_sam_char(ch->data[sp.seq_off + i]) << ((i & 1) ? 0 : 4));
                    }
                    dump(name, sp.l_seq, packed.data(), line);
                }
                if (!ch->error.empty()) out += "#error\t" + ch->error + "\n";
            }
        }
        std::fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    }
    if (cmd == "tag") {
        const size_t n = sizeof kTag / sizeof kTag[0];
        if (argc == 2) { print_help("tag", kTag, n); return 2; }
        Parsed p = parse(argc, argv, 2, "tag", kTag, n);
        exclusive(p, kTag, n, {"kmer-seq", "kmer-file"}, "tag");
        exclusive(p, kTag, n, {"filter-matching", "invert-match"}, "tag");
        exclusive(p, kTag, n, {"q-size", "aho-corasick"}, "tag");
        exclusive(p, kTag, n, {"case-insensitive", "lowercase", "uppercase"}, "tag");
        exclusive(p, kTag, n, {"canonical", "reverse-complement"}, "tag");
        exclusive(p, kTag, n, {"suppress-output", "out-file"}, "tag");
        std::string missing;
        if (!p.has("in-file")) missing += "\n  --in-file <IN_FILE>";
        if (!p.has("kmer-seq") && !p.has("kmer-file")) missing += "\n  <--kmer-seq <KMER_SEQ>...|--kmer-file <KMER_FILE>>";
        if (p.has("suppress-output") && !p.has("out-log") && !p.has("json-log")) missing += "\n  <--out-log [<OUT_LOG>]|--json-log [<JSON_LOG>]>";
        if (!missing.empty()) usage_error("the following required arguments were not provided:" + missing, "tag");
        CmdTag a;
        a.in_file = p.one("in-file");
        a.out_file = opt_str(p, "out-file");
        if (p.has("kmer-seq")) a.kmer_seq = p.values.at("kmer-seq");
        a.kmer_file = opt_str(p, "kmer-file");
        a.reverse_complement = p.has("reverse-complement");
        a.canonical = p.has("canonical");
        if (p.has("tag")) a.tag = p.one("tag");
        a.out_log = opt_str(p, "out-log");
        a.json_log = opt_str(p, "json-log");
        if (p.has("threads")) {
            size_t t = parse_usize(p.one("threads"), "--threads <THREADS>", "tag");
            if (t > 65535) usage_error("invalid value '" + p.one("threads") + "' for '--threads <THREADS>': number too large to fit in target type", "tag");
            a.threads = (int)t;
        }
        a.suppress_output = p.has("suppress-output");
        a.filter_matching = p.has("filter-matching");
        a.invert_match = p.has("invert-match");
        a.case_insensitive = p.has("case-insensitive");
        a.lowercase = p.has("lowercase");
        a.uppercase = p.has("uppercase");
        if (p.has("q-size")) a.q_size = parse_usize(p.one("q-size"), "--q-size <Q_SIZE>", "tag");
        a.aho_corasick = p.has("aho-corasick");
        a.argv = all;
        tag_records(std::move(a));
        return 0;
    }
    std::fprintf(stderr, "error: unrecognized subcommand '%s'\n\nUsage: merkurio <COMMAND>\n\nFor more information, try '--help'.\n", cmd.c_str());
    return 2;
}

}  // namespace

static int guarded_run(int argc, char** argv, FILE* err) {
    try {
        return run(argc, argv);
    } catch (const Error& e) {
        std::fflush(stdout);
        std::fputs(e.report().c_str(), err);
        return 1;
    } catch (const std::exception& e) {
        std::fflush(stdout);
        std::fprintf(err, "Error: %s\n", e.what());
        return 1;
    }
}

// merkurio batch <file>: run several command lines in ONE process (diagnostic, not in the reference).
// Creating a CUDA context costs seconds on a multi-GPU box; the test-suite pays it once per group of
// runs this way. One command per line, fields separated by tabs; leading NAME=VALUE fields are
// environment settings for that command only. The exit status of command i goes to <file>.<i>.rc, what
// it would have printed to stderr to <file>.<i>.err.
static int run_batch(const char* argv0, const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "Error: cannot open %s\n", path.c_str()); return 1; }
    std::string all;
    char buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, n);
    std::fclose(f);
    int index = 0;
    size_t pos = 0;
    while (pos < all.size()) {
        size_t nl = all.find('\n', pos);
        std::string line = all.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = nl == std::string::npos ? all.size() : nl + 1;
        if (line.empty()) continue;
        std::vector<std::string> fields;
        for (size_t q = 0;;) {
            size_t tab = line.find('\t', q);
            fields.push_back(line.substr(q, tab == std::string::npos ? std::string::npos : tab - q));
            if (tab == std::string::npos) break;
            q = tab + 1;
        }
        std::vector<std::string> env_names;
        size_t first = 0;
        for (; first < fields.size(); ++first) {
            const std::string& t = fields[first];
            size_t eq = t.find('=');
            bool is_env = eq != std::string::npos && eq > 0;
            for (size_t i = 0; is_env && i < eq; ++i) is_env = (t[i] >= 'A' && t[i] <= 'Z') || (t[i] >= '0' && t[i] <= '9') || t[i] == '_';
            if (!is_env) break;
            setenv(t.substr(0, eq).c_str(), t.substr(eq + 1).c_str(), 1);
            env_names.push_back(t.substr(0, eq));
        }
        std::vector<char*> av;
        av.push_back(const_cast<char*>(argv0));
        for (size_t i = first; i < fields.size(); ++i) av.push_back(const_cast<char*>(fields[i].c_str()));
        const std::string stem = path + "." + std::to_string(index++);
        std::FILE* err = std::fopen((stem + ".err").c_str(), "wb");
        int rc = guarded_run((int)av.size(), av.data(), err ? err : stderr);
        if (err) std::fclose(err);
        if (std::FILE* r = std::fopen((stem + ".rc").c_str(), "wb")) { std::fprintf(r, "%d\n", rc); std::fclose(r); }
        for (auto& name : env_names) unsetenv(name.c_str());
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 3 && std::strcmp(argv[1], "batch") == 0) return run_batch(argv[0], argv[2]);
    // One command per process: the engines are not torn down piece by piece (EngineSet, device.h) and the
    // process leaves without the CUDA runtime's exit handlers — every output has been closed by then, the
    // driver releases the context with the process. That is 0.2-0.5 s of a run that computes for 0.5 s.
    mkh::g_leave_engines_to_process_exit = !std::getenv("MERKURIO_FULL_TEARDOWN");
    const int rc = guarded_run(argc, argv, stderr);
    if (mkh::g_leave_engines_to_process_exit) {
        std::fflush(nullptr);
        std::cout.flush();
        std::cerr.flush();
        mkh::report_process_time_if_asked();
        _exit(rc);
    }
    return rc;
}
