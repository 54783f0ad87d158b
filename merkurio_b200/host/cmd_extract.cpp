// `extract`: same flags, outputs, logs and counters as the reference's extract_records
// (src/cmd_extract.rs:143-717); the per-record matcher calls are replaced by batches on the GPU.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

#include "commands.h"
#include "device.h"
#include "fasta_pipeline.h"
#include "fastq_pipeline.h"
#include "fastq_stream.h"
#include "helpers.h"
#include "io.h"
#include "logger.h"

namespace mkh {

namespace {

struct OutFile {
    FILE* f = nullptr;
    bool owned = false;
    std::string buf;
    void open(const std::optional<std::string>& path, const char* what) {
        if (!path) { f = stdout; return; }
        f = std::fopen(path->c_str(), "wb");
        if (!f) throw Error("No such file or directory (os error 2)").with_context(std::string(what) + rust_debug_string(*path));
        owned = true;
    }
    void write(const FastxRecord& r) {
        r.write(&buf);
        if (buf.size() >= (1u << 20)) flush();
    }
    void write_raw(const char* p, size_t n) {
        if (n >= (256u << 10)) {  // a long run of records: straight to the file, not through the buffer
            flush();
            if (f) write_all(f, p, n);
            return;
        }
        buf.append(p, n);
        if (buf.size() >= (1u << 20)) flush();
    }
    void flush() {  // throws on a write error (full disk, closed pipe): see write_all
        std::string pending;
        pending.swap(buf);  // the bytes count as handed over even if the write fails: the destructor must not try again
        if (f && !pending.empty()) write_all(f, pending.data(), pending.size());
        if (f) flush_checked(f);
    }
    void close() {  // end of a successful run: the last bytes and the close itself are checked
        flush();
        if (f && owned) { FILE* g = f; f = nullptr; close_checked(g); }
    }
    ~OutFile() {
        try { flush(); } catch (...) {}
        if (f && owned) std::fclose(f);
    }
};

std::string join(const std::vector<std::string>& v, const char* sep) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) { if (i) s += sep; s += v[i]; }
    return s;
}

FastxRecord to_record(const RecMeta& m) {
    FastxRecord r;
    r.id = m.a; r.raw = m.b; r.qual = m.c; r.fastq = m.fastq; r.crlf = m.crlf;
    return r;
}

// Records of the chunked FASTQ ingest carry a (chunk, span) reference instead of strings.
std::string_view record_id(const RecMeta& m) {
    if (m.kind != 1) return m.a;
    const Chunk* ch = static_cast<const Chunk*>(m.chunk);
    const RecSpan& r = ch->recs[m.idx];
    return std::string_view(ch->id(r), r.id_len);
}

void write_record(OutFile& w, const RecMeta& m) {
    if (m.kind == 2) {  // FASTA pipeline: '>' id, the wrapped lines as they are in the file, a final line break
        const FaRecord* rec = static_cast<const FaRecord*>(m.chunk);
        const char* le = rec->crlf ? "\r\n" : "\n";
        w.write_raw(">", 1);
        w.write_raw(rec->id.data(), rec->id.size());
        w.write_raw(le, std::strlen(le));
        for (size_t i = 0; i < rec->raw.size(); ++i) {
            const FaRecord::Range& r = rec->raw[i];
            if (i) w.write_raw("\n", 1);  // chunks are cut at line breaks: exactly one '\n' lies between two ranges
            uint32_t n = r.len;
            if (i + 1 == rec->raw.size() && n && r.chunk->data[r.off + n - 1] == '\r') --n;  // the last line's '\r' is not part of raw_seq
            w.write_raw(r.chunk->data.data() + r.off, n);
        }
        w.write_raw(le, std::strlen(le));
        return;
    }
    if (m.kind != 1) { w.write(to_record(m)); return; }
    const Chunk* ch = static_cast<const Chunk*>(m.chunk);
    const RecSpan& r = ch->recs[m.idx];
    if (r.plain) { w.write_raw(ch->data.data() + r.start, r.end - r.start); return; }
    FastxRecord fr;
    fr.id.assign(ch->id(r), r.id_len);
    fr.raw.assign(ch->seq(r), r.seq_len);
    fr.qual.assign(ch->qual(r), r.seq_len);
    fr.fastq = true;
    fr.crlf = r.crlf;
    w.write(fr);
}

}  // namespace

void extract_records(CmdExtract args) {
    check_log_flag_conflict(args.out_log, args.json_log, args.out_fastx, args.suppress_output);

    std::vector<std::string> pattern_list;
    try {
        pattern_list = parse_pattern_list(args.kmer_file, args.kmer_seq, args.reverse_complement, args.canonical, args.lowercase, args.uppercase);
    } catch (const Error& e) {
        throw e.with_context("Problem parsing pattern list.");
    }
    // which algorithm the reference would run: decides report order, count semantics and the JSON string
    args.aho_corasick = choose_aho_corasick(pattern_list, args.case_insensitive, args.q_size, args.aho_corasick);

    std::unique_ptr<Sink> log_sink;
    if (args.out_log) log_sink = Sink::open(*args.out_log, "Problem creating log file");

    error_if_directory(args.in_fastx, "Record file path");
    const std::string f1 = path_file_name(args.in_fastx);
    std::string f2;
    if (args.in_fastq_2) {
        error_if_directory(*args.in_fastq_2, "Second read file path");
        f2 = path_file_name(*args.in_fastq_2);
    }
    const bool paired = args.in_fastq_2.has_value();
    // what the readers open: the inputs themselves, or seekable copies of inputs that can be read only once
    const std::string in1 = spool_if_not_seekable(args.in_fastx);
    const std::string in2 = paired ? spool_if_not_seekable(*args.in_fastq_2) : std::string();
    const bool logging_active = log_sink || args.json_log;
    BufferedLogger logger(std::move(log_sink), 8192);
    std::unique_ptr<JsonLogger> jl;
    if (args.json_log) jl.reset(new JsonLogger(Sink::open(*args.json_log, "Error creating JSON log file"), 8192));

    if (logging_active) {
        logger.write_header("#SeqKatcher extract log\n");
        logger.write_header("#" + timestamp_now() + "\n");
        logger.write_header(std::string("#Running ") + kProgram + " version " + kVersion + "\n");
        logger.write_header("#Command line: " + join(args.argv, " ") + "\n");
        logger.write_header("#Searching for " + std::to_string(pattern_list.size()) + " pattern" + (pattern_list.size() > 1 ? "s" : "") + " " +
                            (args.invert_match ? "(inverted matching)" : "") + "\n");
        logger.write_header("#\n#File\tRecord\tPattern\tPosition (zero-based)\n");
        logger.flush();
    }
    if (!args.aho_corasick) validate_bndmq(pattern_list, args.q_size);

    std::unique_ptr<FastxReader> reader, reader2;
    try {
        reader.reset(new FastxReader(in1));
    } catch (const Error& e) {
        throw e.with_context("Invalid FASTQ/A input path or file: " + rust_debug_string(args.in_fastx));
    }

    uint64_t nb_records_tot = 0, nb_bases = 0, nb_records_extracted = 0;
    uint64_t nb_hits_tot[2] = {0, 0}, nb_records_hit[2] = {0, 0};
    std::vector<uint64_t> pattern_hit_counts(pattern_list.size(), 0);

    OutFile writer, writer2;
    if (!paired) {
        std::optional<std::string> path;
        if (args.out_fastx) path = path_with_extension(*args.out_fastx, identify_uncompressed_type(args.in_fastx));
        writer.open(path, "Error writing to output file; no such directory: ");
    } else {
        try {
            reader2.reset(new FastxReader(in2));
        } catch (const Error& e) {
            throw e.with_context("Invalid second FASTQ input path or file: Some(" + rust_debug_string(*args.in_fastq_2) + ")");
        }
        std::optional<std::string> p1, p2;
        if (args.out_fastx) {
            std::string base = path_with_extension(*args.out_fastx, identify_uncompressed_type(args.in_fastx));
            p1 = add_suffix_to_file_prefix(base, "_1");
            p2 = add_suffix_to_file_prefix(base, "_2");
        }
        writer.open(p1, "Error writing to paired-end file; no such directory: ");
        writer2.open(p2, "Error writing second paired-end file; no such directory: ");
    }

    auto emit = [&](const std::string& fname, const RecMeta& m, const RecHit& h) {
        const std::string_view id = record_id(m);
        logger.log_fields(fname, id, pattern_list[h.pattern], h.start);
        if (jl) jl->log_fields(fname, id, pattern_list[h.pattern], h.start);
    };
    auto by_pattern_then_start = [](const RecHit& x, const RecHit& y) { return x.pattern != y.pattern ? x.pattern < y.pattern : x.start < y.start; };

    // ---- per-record consumer (src/cmd_extract.rs:321-406) and per-pair consumer (:463-607) --------
    // the FASTQ pipeline adds the totals of a whole batch at once and calls the consumers below only
    // for the records that can produce output
    bool bulk_totals = false;
    RecMeta mate1;
    bool mate1_found = false;
    std::vector<RecHit> mate1_hits;

    auto on_single = [&](RecMeta& m, bool found, std::vector<RecHit>& hits) {
        bool found_occ = false;
        if (logging_active) {
            if (!bulk_totals) {
                nb_records_tot += 1;
                nb_bases += m.len;
            }
            if (args.aho_corasick) {
                for (const RecHit& h : hits) {
                    emit(f1, m, h);
                    pattern_hit_counts[h.pattern] += 1;
                    nb_hits_tot[0] += 1;
                    found_occ = true;
                }
            } else {
                std::stable_sort(hits.begin(), hits.end(), by_pattern_then_start);
                for (size_t i = 0; i < hits.size(); ++i) {
                    emit(f1, m, hits[i]);
                    nb_hits_tot[0] += 1;
                    if (i == 0 || hits[i - 1].pattern != hits[i].pattern) pattern_hit_counts[hits[i].pattern] += 1;
                    found_occ = true;
                }
            }
            if (found_occ) nb_records_hit[0] += 1;
        } else {
            found_occ = found;
        }
        if (found_occ != args.invert_match) {
            nb_records_extracted += 1;
            if (!args.suppress_output) write_record(writer, m);
        }
    };

    auto on_paired = [&](RecMeta& m, bool found, std::vector<RecHit>& hits) {
        if (m.file == 0) {
            mate1 = std::move(m);
            mate1_found = found;
            mate1_hits = hits;
            return;
        }
        RecMeta& m2 = m;
        bool found_occ = false;
        if (logging_active) {
            if (!bulk_totals) {
                nb_records_tot += 2;
                nb_bases += mate1.len + m2.len;
            }
            if (args.aho_corasick) {
                for (const RecHit& h : mate1_hits) { emit(f1, mate1, h); pattern_hit_counts[h.pattern] += 1; nb_hits_tot[0] += 1; }
                for (const RecHit& h : hits) { emit(f2, m2, h); pattern_hit_counts[h.pattern] += 1; nb_hits_tot[1] += 1; }
            } else {
                // for each pattern: its positions in mate 1, then its positions in mate 2 (:543-585)
                std::stable_sort(mate1_hits.begin(), mate1_hits.end(), by_pattern_then_start);
                std::stable_sort(hits.begin(), hits.end(), by_pattern_then_start);
                size_t i = 0, j = 0;
                while (i < mate1_hits.size() || j < hits.size()) {
                    uint32_t p = UINT32_MAX;
                    if (i < mate1_hits.size()) p = mate1_hits[i].pattern;
                    if (j < hits.size()) p = std::min(p, hits[j].pattern);
                    bool any1 = false, any2 = false;
                    for (; i < mate1_hits.size() && mate1_hits[i].pattern == p; ++i) { emit(f1, mate1, mate1_hits[i]); nb_hits_tot[0] += 1; any1 = true; }
                    for (; j < hits.size() && hits[j].pattern == p; ++j) { emit(f2, m2, hits[j]); nb_hits_tot[1] += 1; any2 = true; }
                    pattern_hit_counts[p] += (any1 ? 1 : 0) + (any2 ? 1 : 0);
                }
            }
            if (!mate1_hits.empty()) nb_records_hit[0] += 1;
            if (!hits.empty()) nb_records_hit[1] += 1;
            found_occ = !mate1_hits.empty() || !hits.empty();
        } else {
            found_occ = mate1_found || found;
        }
        if (found_occ != args.invert_match) {
            nb_records_extracted += 2;
            if (!args.suppress_output) {
                write_record(writer, mate1);
                write_record(writer2, m2);
            }
        }
    };

    {
        RecordCallback cb;
        if (paired) cb = on_paired; else cb = on_single;
        const mk_mode mode = logging_active ? MK_MODE_ALL_HITS : MK_MODE_FLAG;
        // 4-line FASTQ (plain or gzip) goes through the reader -> packer -> GPU pipeline (fastq_pipeline.h);
        // this thread then only looks at the records the device flagged. The readers start first: the
        // input is read and indexed while CUDA starts up.
        const bool pipelined = !std::getenv("MERKURIO_NO_FASTQ_PIPELINE") && looks_like_fastq(in1) &&
                               (!paired || looks_like_fastq(in2));
        std::unique_ptr<FastqChunkReader> chunks1, chunks2;
        if (pipelined) {
            chunks1 = FastqPipeline::open_reader(in1, paired ? 2 : 1);
            if (paired) chunks2 = FastqPipeline::open_reader(in2, 2);
        }
        // FASTA (one file, or two files of mates) has its own pipeline (fasta_pipeline.h): records of any length, cut
        // into pieces
        const bool fasta_pipelined = !pipelined && !std::getenv("MERKURIO_NO_FASTA_PIPELINE") && looks_like_fasta(in1) &&
                                     (!paired || looks_like_fasta(in2));
        std::unique_ptr<FastaChunkReader> fa_chunks, fa_chunks2;
        if (fasta_pipelined) {
            fa_chunks = FastaPipeline::open_reader(in1, paired ? 2 : 1);
            if (paired) fa_chunks2 = FastaPipeline::open_reader(in2, 2);
        }
        // the record-by-record path (FASTA, and whatever the FASTQ pipeline does not take) reads ahead on its own
        // threads, also started before the engines
        const bool keep_text = !args.suppress_output;
        std::unique_ptr<PrefetchingFastxReader> ahead1, ahead2;
        if (!pipelined && !fasta_pipelined) {
            reader->set_keep_raw(keep_text);
            ahead1.reset(new PrefetchingFastxReader(reader.get()));
            if (paired) {
                reader2->set_keep_raw(keep_text);
                ahead2.reset(new PrefetchingFastxReader(reader2.get()));
            }
        }
        EngineSet engines(pattern_list, args.case_insensitive, pipelined ? 16 : (fasta_pipelined ? 32 : 64));
        Scanner scanner(engines, MK_ENC_ASCII, mode, cb);
        auto feed = [&](FastxRecord& rec, uint8_t file) {
            RecMeta m;
            m.a = std::move(rec.id);
            if (keep_text) { m.b = std::move(rec.raw); m.c = std::move(rec.qual); }
            m.file = file; m.fastq = rec.fastq; m.crlf = rec.crlf;
            scanner.add_record(rec.seq.data(), rec.seq.size(), std::move(m));
        };
        auto consume_batch = [&](const PackedBatch& b, const mk_result& res) {
            const uint32_t F = paired ? 2 : 1;
            const uint32_t n_units = b.n_records / F;  // records, or pairs
            if (logging_active) {
                nb_records_tot += b.n_records;
                nb_bases += b.n_units;
            }
            size_t cursor[2] = {0, 0}, hi = 0;
            std::vector<RecHit> hits;
            auto deliver = [&](uint32_t u) {
                for (uint32_t f = 0; f < F; ++f) {
                    const uint32_t r = u * F + f;
                    const BatchSeg& sg = b.locate((int)f, u, &cursor[f]);
                    RecMeta m;
                    m.kind = 1;
                    m.chunk = sg.chunk.get();
                    m.idx = sg.first + (u - sg.rec0);
                    const RecSpan& sp = static_cast<const Chunk*>(sg.chunk.get())->recs[m.idx];
                    m.file = (uint8_t)f; m.fastq = true; m.crlf = sp.crlf; m.len = sp.seq_len;
                    hits.clear();
                    while (hi < res.n_hits && res.hits[hi].record < r) ++hi;
                    for (; hi < res.n_hits && res.hits[hi].record == r; ++hi)
                        hits.push_back(RecHit{res.hits[hi].start, res.hits[hi].pattern, res.hits[hi].len});
                    const bool found = (res.record_flags[r >> 6] >> (r & 63)) & 1;
                    cb(m, found, hits);
                }
            };
            if (args.invert_match && !logging_active && !paired && !args.suppress_output) {
                // -v without logs writes nearly every record: runs of unflagged records that lie back to back in one
                // chunk, already in the writer's form, are copied in one piece (what on_single would do one by one)
                auto flagged = [&](uint32_t r) { return (res.record_flags[r >> 6] >> (r & 63)) & 1; };
                size_t at = 0;
                for (uint32_t u = 0; u < n_units;) {
                    if (flagged(u)) { ++u; continue; }
                    const BatchSeg& sg = b.locate(0, u, &at);
                    const Chunk* ch = static_cast<const Chunk*>(sg.chunk.get());
                    const RecSpan& first = ch->recs[sg.first + (u - sg.rec0)];
                    if (!first.plain) { deliver(u); ++u; continue; }
                    uint32_t end = first.end, v = u + 1;
                    for (const uint32_t seg_end = sg.rec0 + sg.count; v < seg_end && !flagged(v); ++v) {
                        const RecSpan& next = ch->recs[sg.first + (v - sg.rec0)];
                        if (!next.plain || next.start != end) break;
                        end = next.end;
                    }
                    writer.write_raw(ch->data.data() + first.start, end - first.start);
                    nb_records_extracted += v - u;
                    u = v;
                }
                return;
            }
            if (args.invert_match) {
                for (uint32_t u = 0; u < n_units; ++u) deliver(u);
                return;
            }
            const size_t words = ((size_t)b.n_records + 63) / 64;
            for (size_t w = 0; w < words; ++w) {
                uint64_t x = res.record_flags[w];
                if (paired) x = (x | (x >> 1)) & 0x5555555555555555ull;  // a pair is flagged through either mate
                while (x) {
                    const uint32_t r = (uint32_t)(w * 64 + (size_t)__builtin_ctzll(x));
                    x &= x - 1;
                    if (r < b.n_records) deliver(r / F);
                }
            }
        };
        // FASTA pipeline: assemble the records from their pieces; every record goes through the per-record consumer
        auto consume_fasta = [&](const PackedBatch& b, const mk_result& res) {
            const FaBatchInfo& fi = FastaPipeline::info(b);
            size_t hi = 0;
            for (uint32_t r = 0; r < b.n_records; ++r) {
                const FaPiece& pc = fi.pieces[r];
                FaRecord& rec = *pc.rec;
                if ((res.record_flags[r >> 6] >> (r & 63)) & 1) rec.found = true;
                for (; hi < res.n_hits && res.hits[hi].record == r; ++hi) {
                    const mk_hit& h = res.hits[hi];
                    if (!pc.first && (uint64_t)h.start + h.len <= pc.own_from) continue;  // owned by the previous piece
                    rec.hits.push_back(RecHit{pc.base + h.start, h.pattern, h.len});
                }
                if (!pc.last) continue;
                RecMeta m;
                m.kind = 2;
                m.chunk = &rec;
                m.keep = pc.rec;  // a first mate waits for the second one, which may end in a later batch
                m.file = rec.file;
                m.a = rec.id;
                m.len = (uint32_t)rec.len;
                m.crlf = rec.crlf;
                cb(m, rec.found, rec.hits);
            }
        };
        FastxRecord r1, r2;
        try {
            if (fasta_pipelined) {
                reader.reset();
                reader2.reset();
                FastaPipeline pipe(engines, std::move(fa_chunks), std::move(fa_chunks2), mode, keep_text, consume_fasta);
                pipe.run();
            } else if (pipelined) {
                reader.reset();
                reader2.reset();
                bulk_totals = true;
                FastqPipeline pipe(engines, std::move(chunks1), std::move(chunks2), mode, consume_batch);
                pipe.run();
            } else if (!paired) {
                for (;;) {
                    bool more;
                    try { more = ahead1->next(&r1); } catch (const Error& e) { throw e.with_context("Error during FASTQ/A record parsing."); }
                    if (!more) break;
                    feed(r1, 0);
                }
            } else {
                for (;;) {
                    bool more;
                    try { more = ahead1->next(&r1); } catch (const Error& e) { throw e.with_context("Error during FASTQ record parsing of first file."); }
                    if (!more) break;
                    bool more2;
                    try { more2 = ahead2->next(&r2); } catch (const Error& e) {
                        throw e.with_context("Error during FASTQ record parsing of second file. Do the two input files contain the same number of records?");
                    }
                    if (!more2) throw Error("Error during FASTQ record parsing of second file. Do the two input files contain the same number of records?");
                    feed(r1, 0);
                    feed(r2, 1);
                }
                if (ahead2->next(&r2))
                    throw Error("The two input files have a different number of records. Please provide valid paired-end read files.");
            }
        } catch (...) {
            // the reference has already written everything up to the failing record
            scanner.finish();
            writer.flush(); writer2.flush(); logger.flush();
            if (jl) jl->flush();
            throw;
        }
        scanner.finish();
    }
    writer.close();
    writer2.close();
    const double t_sum0 = steady_seconds();

    size_t nb_patterns_found = 0;
    for (uint64_t c : pattern_hit_counts) nb_patterns_found += c > 0;
    if (logging_active) {
        logger.flush();
        char pct[64];
        std::snprintf(pct, sizeof pct, "%.2f", (double)nb_patterns_found / (double)pattern_hit_counts.size() * 100.0);
        logger.write_header("#\n#Number of patterns found: " + std::to_string(nb_patterns_found) + "/" + std::to_string(pattern_hit_counts.size()) + " (" + pct + " %)\n");
        logger.write_header("#Pattern\tCount\n");
        if (logger.active()) {  // a million queries: one write per MiB, not per line
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
        logger.write_header("#Total number of hits: " + std::to_string(nb_hits_tot[0] + nb_hits_tot[1]) + "\n");
        logger.write_header("#Number of distinct records with a hit: " + std::to_string(nb_records_hit[0] + nb_records_hit[1]) + "\n");
        if (paired) {
            logger.write_header("#\n#Total number of hits in file 1: " + std::to_string(nb_hits_tot[0]) + "\n");
            logger.write_header("#Total number of hits in file 2: " + std::to_string(nb_hits_tot[1]) + "\n");
            logger.write_header("#Number of distinct records with a hit in file 1: " + std::to_string(nb_records_hit[0]) + "\n");
            logger.write_header("#Number of distinct records with a hit in file 2: " + std::to_string(nb_records_hit[1]) + "\n");
            logger.write_header("#Total number of extracted records: " + std::to_string(nb_records_extracted) + "\n");
        }
        logger.flush();
    }
    if (jl) {
        Json input_files = Json::object();
        input_files["kmer_file"] = args.kmer_file ? Json::string(*args.kmer_file) : Json::null();
        input_files["record_file_1"] = Json::string(f1);
        input_files["record_file_2"] = paired ? Json::string(f2) : Json::null();
        Json cmdline = Json::array();
        for (auto& a : args.argv) cmdline.a.push_back(Json::string(a));
        Json meta = Json::object();
        meta["program"] = Json::string(kProgram);
        meta["version"] = Json::string(kVersion);
        meta["timestamp"] = Json::string(timestamp_now());
        meta["subcommand"] = Json::string("extract");
        meta["command_line"] = cmdline;
        meta["search_algorithm"] = Json::string(args.aho_corasick ? "Aho-Corasick" : "BNDMq");
        meta["inverted_matching"] = Json::boolean(args.invert_match);
        meta["case_insensitive"] = Json::boolean(args.case_insensitive);
        meta["input_files"] = input_files;
        Json summary = Json::object();
        summary["number_of_patterns_searched"] = Json::integer((int64_t)pattern_list.size());
        summary["number_of_patterns_found"] = Json::integer((int64_t)nb_patterns_found);
        summary["number_of_records_searched"] = Json::integer((int64_t)nb_records_tot);
        summary["number_of_characters_searched"] = Json::integer((int64_t)nb_bases);
        summary["number_of_matches"] = Json::integer((int64_t)(nb_hits_tot[0] + nb_hits_tot[1]));
        summary["number_of_distinct_records_with_a_hit"] = Json::integer((int64_t)(nb_records_hit[0] + nb_records_hit[1]));
        Json pstats = Json::object();
        pstats["searching_paired_end_reads"] = Json::boolean(paired);
        pstats["number_of_hits_in_file_1"] = Json::integer((int64_t)nb_hits_tot[0]);
        pstats["number_of_hits_in_file_2"] = paired ? Json::integer((int64_t)nb_hits_tot[1]) : Json::null();
        pstats["number_of_distinct_records_with_a_hit_in_file_1"] = Json::integer((int64_t)nb_records_hit[0]);
        pstats["number_of_distinct_records_with_a_hit_in_file_2"] = paired ? Json::integer((int64_t)nb_records_hit[1]) : Json::null();
        pstats["number_of_extracted_records"] = Json::integer((int64_t)nb_records_extracted);
        jl->finalize(meta, pattern_list, pattern_hit_counts, summary, &pstats);
    }
    if (std::getenv("MERKURIO_TIMING")) std::fprintf(stderr, "[merkurio] summaries of the logs: %.3f s\n", steady_seconds() - t_sum0);
}

}  // namespace mkh
