// Input side: a buffered byte source over any supported compression (codecs.h), the FASTA/FASTQ
// record reader (the role needletail plays in src/cmd_extract.rs:281,321,412,463) and the SAM/BAM
// record reader (the role of the `bam` crate in src/cmd_tag.rs:504-613).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "codecs.h"
#include "common.h"

namespace mkh {

// Threads used to inflate BGZF input (BAM, bgzip'ed text). Default: half the cores, at most 8; `tag -p N`
// raises it to at least N. 1 = plain zlib stream.
void set_decompression_threads(int n);
int decompression_threads();

// Inputs are opened more than once (format sniffing by content, then the reader the format calls for; BGZF detection
// reads the head of the file with pread). A FIFO, a process substitution (`-i <(zcat x.fq.gz)`) or /dev/stdin can be
// read only once: such an input is copied into an unlinked temporary file first and the path of that copy
// (/proc/self/fd/N) is returned; regular files and paths that cannot be opened come back unchanged (the reader then
// raises the usual error). The reference opens its input once and sniffs from the stream (needletail).
std::string spool_if_not_seekable(const std::string& path);

class ByteSource {
public:
    explicit ByteSource(const std::string& path);
    ~ByteSource();
    ByteSource(const ByteSource&) = delete;
    // Reads one line (without the '\n'; a trailing '\r' is kept). Returns false at end of input.
    bool getline(std::string* line);
    // Same, without a copy when the whole line is already buffered: *p / *n then point into the buffer
    // (valid until the next call); otherwise into *spill.
    bool getline_view(const char** p, size_t* n, std::string* spill);
    // Reads exactly n bytes; returns false on EOF before the first byte, throws on a short read.
    bool read_exact(void* dst, size_t n);
    // Reads up to n bytes (what is buffered first); 0 at end of input.
    size_t read_some(void* dst, size_t n);
    int peek();  // next byte or -1
private:
    bool fill();
    std::unique_ptr<InputStream> in_;  // plain, gzip, BGZF (block-parallel), bzip2 or xz
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
};

// One FASTA/FASTQ record as needletail exposes it.
struct FastxRecord {
    std::string id;    // header line without '>' / '@' (and without the line break)
    std::string seq;   // line breaks removed (record.seq())
    std::string raw;   // FASTA: sequence lines as in the file, last line break dropped; FASTQ: == seq
    std::string qual;  // FASTQ only
    bool fastq = false;
    bool crlf = false;
    // record.write(writer, None): FASTA keeps its wrapping; FASTQ is four lines with a bare '+'
    void write(std::string* out) const;
};

class FastxReader {
public:
    explicit FastxReader(const std::string& path);
    bool next(FastxRecord* rec);  // throws Error on malformed input
    // FASTA: do not keep the wrapped text of the records (FastxRecord::raw stays empty) when nothing will be written
    void set_keep_raw(bool keep) { keep_raw_ = keep; }
private:
    ByteSource src_;
    std::string pending_;  // a header line already consumed
    bool have_pending_ = false, started_ = false, fastq_ = false, keep_raw_ = true;
};

// Runs a FastxReader on its own thread, ahead of the consumer by up to `budget_bytes` of record text
// (FASTA input: chromosomes are read and unwrapped while CUDA starts up and while earlier ones are
// scanned). Same results and errors, raised at the same record, as calling FastxReader::next directly.
class PrefetchingFastxReader {
public:
    explicit PrefetchingFastxReader(FastxReader* reader, size_t budget_bytes = (size_t)1 << 30);
    ~PrefetchingFastxReader();
    bool next(FastxRecord* rec);

private:
    struct State;
    std::unique_ptr<State> st_;
};

// One alignment record. `packed` holds the sequence as BAM stores it (4 bits per base, first base in
// the high nibble, "=ACMGRSVTWYHKDBN"); the device scans it directly.
struct AlnRecord {
    std::string name;
    std::string sam_line;          // the record as a SAM text line (no line break)
    std::vector<uint8_t> packed;   // (l_seq + 1) / 2 bytes
    uint32_t l_seq = 0;
};

class AlnReader {
public:
    AlnReader(const std::string& path, bool is_bam);
    const std::vector<std::string>& header_lines() const { return header_; }
    const std::vector<std::string>& refs() const { return refs_; }
    bool is_bam() const { return bam_; }
    ByteSource& source() { return src_; }  // positioned after the header: the chunked reader continues from here
    bool next(AlnRecord* rec);
private:
    void read_bam_header();
    bool next_sam(AlnRecord* rec);
    bool next_bam(AlnRecord* rec);
    ByteSource src_;
    bool bam_;
    std::vector<std::string> header_, refs_;
    std::string pending_;
    bool have_pending_ = false;
};

extern const char kNibbleChars[17];

// BAM code of a SAM sequence character: case-insensitive, anything outside "=ACMGRSVTWYHKDBN" is N.
uint8_t nibble_of_sam_char(char c);

// One BAM alignment (the `block_size` bytes after the length field) as a SAM text line without line
// break. Throws Error on a malformed record. name: the read name; packed / l_seq: the 4-bit sequence.
void bam_body_to_sam(const char* body, size_t len, const std::vector<std::string>& refs, std::string* name, std::string* line,
                     const uint8_t** packed, uint32_t* l_seq);

}  // namespace mkh
