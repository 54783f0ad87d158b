// C ABI of the host's input streams (include/merkurio_io.h): one decompressed byte stream over a FASTA / FASTQ / SAM
// input whatever its compression — plain, gzip (one member on several threads: pgzip.cpp), BGZF (block-parallel), bzip2,
// xz, zstd — for a host in another language: the reference's Rust host reads its inputs through
// needletail::parse_fastx_file (src/cmd_extract.rs:281,412), which decompresses on the calling thread; handed this
// stream as a `Read` (needletail::parse_fastx_reader) it gets the ingest rate of this repository's own CLI.
#include <cstring>
#include <memory>
#include <string>

#include "../../include/merkurio_io.h"
#include "codecs.h"
#include "common.h"

struct mk_input {
    std::unique_ptr<mkh::InputStream> stream;
    std::string error;
};

namespace {
thread_local std::string g_open_error;
}

extern "C" {

mk_input* mk_input_open(const char* path) {
    if (!path) {
        g_open_error = "null path";
        return nullptr;
    }
    try {
        std::unique_ptr<mk_input> in(new mk_input);
        in->stream = mkh::InputStream::open(path);
        return in.release();
    } catch (const std::exception& e) {
        g_open_error = e.what();
        return nullptr;
    }
}

long long mk_input_read(mk_input* in, void* dst, unsigned long long n) {
    if (!in || (!dst && n)) return -1;
    if (!in->error.empty()) return -1;
    try {
        return (long long)in->stream->read(static_cast<char*>(dst), (size_t)n);
    } catch (const std::exception& e) {
        in->error = e.what();
        return -1;
    }
}

const char* mk_input_error(const mk_input* in) { return in ? in->error.c_str() : g_open_error.c_str(); }

void mk_input_close(mk_input* in) { delete in; }

}  // extern "C"
