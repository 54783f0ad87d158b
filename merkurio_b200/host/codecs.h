// One byte stream over an input file, whatever its compression: plain, gzip, BGZF (block-parallel,
// bgzf.h), bzip2 or xz — the formats needletail's `compression` feature opens for the reference
// (README.md:39, manual/src/extract.md:7), recognised by their magic bytes like needletail does, not
// by the file name. zlib is linked; libbz2 and liblzma have no headers in this image and are bound at
// run time (dlopen of libbz2.so.1.0 / liblzma.so.5 with the prototypes of their stable C ABIs).
#pragma once
#include <cstddef>
#include <memory>
#include <string>

namespace mkh {

class InputStream {
public:
    virtual ~InputStream() {}
    // Reads up to n bytes (at least one unless the input is exhausted); 0 at end of input.
    virtual size_t read(char* dst, size_t n) = 0;
    // An uncompressed regular file can be read in parallel: n bytes at the current position, sliced over `threads`
    // concurrent pread calls (a single read(2) copies out of the page cache at about 5 GB/s). Other streams: read().
    virtual size_t read_parallel(char* dst, size_t n, int threads) { (void)threads; return read(dst, n); }
    // Opens `path`; throws Error("No such file or directory (os error 2)") if it cannot be opened.
    static std::unique_ptr<InputStream> open(const std::string& path);
};

}  // namespace mkh
