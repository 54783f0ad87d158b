// `tag`: same flags, outputs, logs and counters as the reference's tag_records
// (src/cmd_tag.rs:155-689). BAM sequences go to the device in their 4-bit packing; SAM text is
// packed to the same codes (the reference sees both through `record.sequence()`, src/cmd_tag.rs:395).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

#include "aln_pipeline.h"
#include "bamwriter.h"
#include "commands.h"
#include "device.h"
#include "helpers.h"
#include "io.h"
#include "logger.h"

namespace mkh {

namespace {

std::string join(const std::vector<std::string>& v, const char* sep) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) { if (i) s += sep; s += v[i]; }
    return s;
}

// existing value of an optional field of a SAM line: 0 = absent, 1 = string (value in *val), 2 = other type
int existing_tag(const std::string& line, const std::string& tag, std::string* val) {
    size_t pos = 0;
    int field = 0;
    while (pos <= line.size()) {
        size_t tab = line.find('\t', pos);
        size_t end = tab == std::string::npos ? line.size() : tab;
        if (field >= 11 && end - pos >= 5 && line.compare(pos, 2, tag) == 0 && line[pos + 2] == ':') {
            if (line[pos + 3] == 'Z' && line[pos + 4] == ':') { *val = line.substr(pos + 5, end - pos - 5); return 1; }
            return 2;
        }
        if (tab == std::string::npos) break;
        pos = tab + 1;
        ++field;
    }
    return 0;
}

bool bytes_less(const std::string& a, const std::string& b) {
    int c = std::memcmp(a.data(), b.data(), std::min(a.size(), b.size()));
    return c != 0 ? c < 0 : a.size() < b.size();
}

}  // namespace

void tag_records(CmdTag args) {
    check_log_flag_conflict(args.out_log, args.json_log, args.out_file, args.suppress_output);
    error_if_directory(args.in_file, "Record file path");
    const std::string in_name = path_file_name(args.in_file);

    std::vector<std::string> pattern_list;
    try {
        pattern_list = parse_pattern_list(args.kmer_file, args.kmer_seq, args.reverse_complement, args.canonical, args.lowercase, args.uppercase);
    } catch (const Error& e) {
        throw e.with_context("Problem parsing pattern list.");
    }
    args.aho_corasick = choose_aho_corasick(pattern_list, args.case_insensitive, args.q_size, args.aho_corasick);

    std::unique_ptr<Sink> log_sink;
    if (args.out_log) log_sink = Sink::open(*args.out_log, "Problem creating log file");
    const bool logging_active = log_sink || args.json_log;
    if (args.threads < 1) throw Error("Number of threads must be at least 1.");
    if (args.tag.size() != 2) throw Error("Tag must be exactly two characters long.");
    if (!args.aho_corasick) validate_bndmq(pattern_list, args.q_size);

    std::string in_ext;
    if (!path_extension(args.in_file, &in_ext)) throw Error("Could not detect the file extension: " + rust_debug_string(args.in_file));
    std::string out_ext = "STDOUT";
    if (args.out_file) {
        if (!path_extension(*args.out_file, &out_ext)) out_ext = in_ext;
    }

    BufferedLogger logger(std::move(log_sink), 8192);
    std::unique_ptr<JsonLogger> jl;
    if (args.json_log) jl.reset(new JsonLogger(Sink::open(*args.json_log, "Error creating JSON log file"), 8192));
    if (logging_active) {
        logger.write_header("#SeqKatcher tag log\n");
        logger.write_header("#" + timestamp_now() + "\n");
        logger.write_header(std::string("#Running ") + kProgram + " version " + kVersion + "\n");
        logger.write_header("#Command line: " + join(args.argv, " ") + "\n");
        logger.write_header("#Tag used for labeling records: " + args.tag + "\n");
        logger.write_header("#Searching for " + std::to_string(pattern_list.size()) + " pattern" + (pattern_list.size() > 1 ? "s" : "") + " " +
                            (args.invert_match ? "(inverted matching)" : "") + "\n");
        logger.write_header("#\n#File\tRecord\tPattern\tPosition (zero-based)\n");
        logger.flush();
    }

    if (in_ext != "bam" && in_ext != "sam") throw Error("Input file must be a BAM or SAM file.");
    // -p: additional (de)compression threads in the reference (src/cmd_tag.rs:504-507); here a lower bound
    set_decompression_threads(std::max(decompression_threads(), args.threads));
    std::unique_ptr<AlnReader> reader;
    try {
        reader.reset(new AlnReader(spool_if_not_seekable(args.in_file), in_ext == "bam"));
    } catch (const Error& e) {
        throw e.with_context(std::string("Error reading ") + (in_ext == "bam" ? "BAM" : "SAM") + " file: " + rust_debug_string(args.in_file));
    }
    std::vector<std::string> header = reader->header_lines();
    header.push_back(std::string("@PG\tID:") + kProgram + "\tPN:" + kProgram + "\tCL:" + join(args.argv, " ") + "\tVN:" + kVersion);
    if (args.suppress_output) header.clear();

    if (out_ext != "bam" && out_ext != "sam" && out_ext != "STDOUT")
        throw Error("Output file must be a BAM or SAM file.").with_context("Could not create writer.");
    std::unique_ptr<BamWriter> bam_out;
    if (out_ext == "bam") {
        try {
            bam_out.reset(new BamWriter(path_with_extension(*args.out_file, "bam"), header, std::max(decompression_threads(), args.threads)));
        } catch (const Error& e) {
            throw e.with_context("Could not create writer.");
        }
    }
    FILE* out = stdout;
    bool out_owned = false;
    if (out_ext == "sam") {
        std::string path = path_with_extension(*args.out_file, "sam");
        out = std::fopen(path.c_str(), "wb");
        if (!out) throw Error("No such file or directory (os error 2)").with_context("Error writing SAM file: " + path).with_context("Could not create writer.");
        out_owned = true;
    }
    std::string obuf;
    auto flush_out = [&] {
        std::string pending;
        pending.swap(obuf);
        write_all(out, pending.data(), pending.size());  // throws: "Error writing record to output file" (src/cmd_tag.rs:494-496)
        flush_checked(out);
    };
    if (!bam_out) for (auto& h : header) { obuf += h; obuf += '\n'; }

    uint64_t nb_records_tot = 0, nb_bases = 0, nb_hits_tot = 0, nb_records_hit = 0;
    std::vector<uint64_t> pattern_hit_counts(pattern_list.size(), 0);
    auto by_pattern_then_start = [](const RecHit& x, const RecHit& y) { return x.pattern != y.pattern ? x.pattern < y.pattern : x.start < y.start; };

    // the SAM/BAM pipeline adds the totals of a whole batch at once
    bool bulk_totals = false;
    // process_record, src/cmd_tag.rs:367-500
    auto on_record = [&](RecMeta& m, bool /*found*/, std::vector<RecHit>& hits) {
        std::vector<std::string> kmers_found;
        if (logging_active) {
            if (!args.aho_corasick) std::stable_sort(hits.begin(), hits.end(), by_pattern_then_start);
            for (size_t i = 0; i < hits.size(); ++i) {
                const RecHit& h = hits[i];
                nb_hits_tot += 1;
                logger.log_fields(in_name, m.a, pattern_list[h.pattern], h.start);
                if (jl) jl->log_fields(in_name, m.a, pattern_list[h.pattern], h.start);
                if (args.aho_corasick) {
                    kmers_found.push_back(pattern_list[h.pattern]);
                    pattern_hit_counts[h.pattern] += 1;
                } else if (i == 0 || hits[i - 1].pattern != h.pattern) {
                    kmers_found.push_back(pattern_list[h.pattern]);
                    pattern_hit_counts[h.pattern] += 1;
                }
            }
            if (!bulk_totals) {
                nb_records_tot += 1;
                nb_bases += m.len;
            }
            if (!kmers_found.empty()) nb_records_hit += 1;
        } else {
            for (const RecHit& h : hits) kmers_found.push_back(pattern_list[h.pattern]);
        }
        bool keep = args.filter_matching ? !kmers_found.empty() : (args.invert_match ? kmers_found.empty() : true);
        if (!keep) return;
        std::string val;
        // BAM -> BAM: the record stays in its binary form (m.chunk / m.idx name it); else its SAM text is in m.b
        const AlnChunk* raw_chunk = static_cast<const AlnChunk*>(m.chunk);
        const AlnSpan* raw = raw_chunk ? &raw_chunk->recs[m.idx] : nullptr;
        int kind = raw ? bam_find_tag(raw_chunk->data.data() + raw->off, raw->len, args.tag, &val) : existing_tag(m.b, args.tag, &val);
        if (kind == 2) throw Error("Invalid tag value format. Expected string value.");
        if (kind == 1 && !val.empty()) {
            size_t pos = 0;
            for (;;) {
                size_t c = val.find(',', pos);
                kmers_found.push_back(val.substr(pos, c == std::string::npos ? std::string::npos : c - pos));
                if (c == std::string::npos) break;
                pos = c + 1;
            }
        }
        std::sort(kmers_found.begin(), kmers_found.end(), bytes_less);
        kmers_found.erase(std::unique(kmers_found.begin(), kmers_found.end()), kmers_found.end());
        if (!args.suppress_output) {
            if (raw) {
                bam_out->write_bam_record(raw_chunk->data.data() + raw->off, raw->len, args.tag, join(kmers_found, ","));
            } else if (bam_out) {
                bam_out->write_sam_line(m.b + "\t" + args.tag + ":Z:" + join(kmers_found, ","));
            } else {
                obuf += m.b; obuf += '\t'; obuf += args.tag; obuf += ":Z:"; obuf += join(kmers_found, ","); obuf += '\n';
                if (obuf.size() >= (1u << 20)) flush_out();
            }
        }
    };

    try {
        const mk_mode mode = logging_active ? MK_MODE_ALL_HITS : MK_MODE_PATTERN_SET;
        const char* kind_name = in_ext == "bam" ? "BAM" : "SAM";
        if (!std::getenv("MERKURIO_NO_ALN_PIPELINE")) {
            // reader -> packer -> GPU pipeline (aln_pipeline.h); the reader starts before the engines so that
            // reading and indexing the input overlaps CUDA start-up
            const size_t chunk_bytes = std::getenv("MERKURIO_CHUNK_BYTES") ? (size_t)std::strtoull(std::getenv("MERKURIO_CHUNK_BYTES"), nullptr, 10)
                                                                           : (size_t)8 << 20;
            std::unique_ptr<AlnChunkReader> chunks(new AlnChunkReader(std::move(reader), chunk_bytes));
            const std::vector<std::string> refs = chunks->refs();
            // BAM in, BAM out, same reference list: kept records are copied as they are, with the tag appended
            const bool passthrough = chunks->is_bam() && bam_out && !args.suppress_output && refs == bam_out->ref_names() &&
                                     !std::getenv("MERKURIO_NO_BAM_PASSTHROUGH");
            EngineSet engines(pattern_list, args.case_insensitive, 16);
            bulk_totals = true;
            auto consume_batch = [&](const PackedBatch& b, const mk_result& res) {
                if (logging_active) {
                    nb_records_tot += b.n_records;
                    nb_bases += b.total_bases;
                }
                size_t cursor = 0, hi = 0;
                std::vector<RecHit> hits;
                RecMeta m;
                auto deliver = [&](uint32_t r) {
                    const BatchSeg& sg = b.locate(0, r, &cursor);
                    const AlnChunk* ch = static_cast<const AlnChunk*>(sg.chunk.get());
                    const AlnSpan& sp = ch->recs[sg.first + (r - sg.rec0)];
                    hits.clear();
                    while (hi < res.n_hits && res.hits[hi].record < r) ++hi;
                    for (; hi < res.n_hits && res.hits[hi].record == r; ++hi)
                        hits.push_back(RecHit{res.hits[hi].start, res.hits[hi].pattern, res.hits[hi].len});
                    m.len = sp.l_seq;
                    m.a.assign(ch->data.data() + sp.name_off, sp.name_len);
                    const bool keep = args.filter_matching ? !hits.empty() : (args.invert_match ? hits.empty() : true);
                    m.b.clear();
                    m.chunk = nullptr;
                    if (keep && passthrough) {
                        m.chunk = ch;
                        m.idx = sg.first + (r - sg.rec0);
                    } else if (keep) {  // the record's SAM text is only needed if it is written
                        if (!ch->bam) {
                            m.b.assign(ch->data.data() + sp.off, sp.len);
                        } else {
                            try {
                                std::string name;
                                bam_body_to_sam(ch->data.data() + sp.off, sp.len, refs, &name, &m.b, nullptr, nullptr);
                            } catch (const Error& e) {
                                throw Error(std::string("Error during BAM record parsing: ") + e.what());
                            }
                        }
                    }
                    on_record(m, !hits.empty(), hits);
                };
                if (!args.filter_matching) {
                    for (uint32_t r = 0; r < b.n_records; ++r) deliver(r);
                    return;
                }
                // keep-only-matching: only the records the device flagged produce output or log lines
                const size_t words = ((size_t)b.n_records + 63) / 64;
                for (size_t w = 0; w < words; ++w) {
                    uint64_t x = res.record_flags[w];
                    while (x) {
                        const uint32_t r = (uint32_t)(w * 64 + (size_t)__builtin_ctzll(x));
                        x &= x - 1;
                        if (r < b.n_records) deliver(r);
                    }
                }
            };
            AlnPipeline pipe(engines, std::move(chunks), mode, consume_batch);
            pipe.run();
        } else {
        EngineSet engines(pattern_list, args.case_insensitive);
        Scanner scanner(engines, MK_ENC_BAM4, mode, on_record);
        AlnRecord rec;
        for (;;) {
            bool more;
            try { more = reader->next(&rec); } catch (const Error& e) {
                scanner.finish();
                throw Error(std::string("Error during ") + kind_name + " record parsing: " + e.what());
            }
            if (!more) break;
            RecMeta m;
            m.a = std::move(rec.name);
            m.b = std::move(rec.sam_line);
            scanner.add_record_packed(rec.packed.data(), rec.l_seq, std::move(m));
        }
        scanner.finish();
        }
    } catch (...) {
        try { flush_out(); } catch (...) {}  // the error on its way out is the one to report
        if (out_owned) std::fclose(out);
        throw;
    }
    try {
        flush_out();
    } catch (...) {
        if (out_owned) std::fclose(out);
        throw;
    }
    if (out_owned) close_checked(out);
    if (bam_out) bam_out->close();

    size_t nb_patterns_found = 0;
    for (uint64_t c : pattern_hit_counts) nb_patterns_found += c > 0;
    if (logging_active) {
        logger.flush();
        char pct[64];
        std::snprintf(pct, sizeof pct, "%.2f", (double)nb_patterns_found / (double)pattern_hit_counts.size() * 100.0);
        logger.write_header("#\n#Number of patterns found: " + std::to_string(nb_patterns_found) + "/" + std::to_string(pattern_hit_counts.size()) + " (" + pct + " %)\n");
        logger.write_header("#Pattern\tCount\n");
        if (logger.active()) {  // large query lists: one write per MiB, not per line
            std::string block;
            for (size_t i = 0; i < pattern_list.size(); ++i) {
                block += '#';
                block += pattern_list[i];
                block += '\t';
                block += std::to_string(pattern_hit_counts[i]);
                block += '\n';
                if (block.size() >= (1u << 20)) { logger.write_header(block); block.clear(); }
            }
            if (!block.empty()) logger.write_header(block);
        }
        logger.write_header("#\n#Total number of records searched: " + std::to_string(nb_records_tot) + "\n");
        logger.write_header("#Total number of characters searched: " + std::to_string(nb_bases) + "\n");
        logger.write_header("#Total number of hits: " + std::to_string(nb_hits_tot) + "\n");
        logger.write_header("#Number of distinct records with a hit: " + std::to_string(nb_records_hit) + "\n");
        logger.flush();
    }
    if (jl) {
        Json input_files = Json::object();
        input_files["kmer_file"] = args.kmer_file ? Json::string(*args.kmer_file) : Json::null();
        input_files["record_file_1"] = Json::string(in_name);
        Json cmdline = Json::array();
        for (auto& a : args.argv) cmdline.a.push_back(Json::string(a));
        Json meta = Json::object();
        meta["program"] = Json::string(kProgram);
        meta["version"] = Json::string(kVersion);
        meta["timestamp"] = Json::string(timestamp_now());
        meta["subcommand"] = Json::string("tag");
        meta["command_line"] = cmdline;
        meta["search_algorithm"] = Json::string(args.aho_corasick ? "Aho-Corasick" : "BNDMq");
        meta["inverted_matching"] = Json::boolean(args.invert_match);
        meta["case_insensitive"] = Json::boolean(args.case_insensitive);
        meta["input_files"] = input_files;
        meta["tag"] = Json::string(args.tag);
        Json summary = Json::object();
        summary["number_of_patterns_searched"] = Json::integer((int64_t)pattern_list.size());
        summary["number_of_patterns_found"] = Json::integer((int64_t)nb_patterns_found);
        summary["number_of_records_searched"] = Json::integer((int64_t)nb_records_tot);
        summary["number_of_characters_searched"] = Json::integer((int64_t)nb_bases);
        summary["number_of_matches"] = Json::integer((int64_t)nb_hits_tot);
        summary["number_of_distinct_records_with_a_hit"] = Json::integer((int64_t)nb_records_hit);
        jl->finalize(meta, pattern_list, pattern_hit_counts, summary, nullptr);
    }
}

}  // namespace mkh
