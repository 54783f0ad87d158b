/* merkurio_io.h — C ABI of the host-side input streams (libmerkurio_io.so; no CUDA, no GPU needed).
 *
 * Replaces, for a host that wants it, the decompression inside needletail::parse_fastx_file
 * (/root/reference/src/cmd_extract.rs:281 and :412 open the inputs with it; needletail's "compression" feature,
 * Cargo.toml:26, inflates on the calling thread through flate2): one byte stream over an input file whatever its
 * compression, recognised by its magic bytes — plain, gzip (a single member is decoded by several threads:
 * merkurio_b200/host/pgzip.cpp), BGZF (block-parallel), bzip2, xz, zstd. The Rust host wraps the handle in a type that
 * implements std::io::Read and hands it to needletail::parse_fastx_reader (INTEGRATION.md has the dozen lines).
 *
 * Errors: mk_input_open returns NULL, mk_input_read returns -1; mk_input_error(handle or NULL) has the text — the
 * same texts the C++ host prints ("No such file or directory (os error 2)", "Error while decompressing the input
 * (truncated gzip stream)", ...). Bytes decoded in front of a damaged spot are handed out before the error.
 * A handle belongs to one thread at a time.
 */
#ifndef MERKURIO_IO_H
#define MERKURIO_IO_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mk_input mk_input;

/* Opens `path`. NULL on failure (mk_input_error(NULL) on the same thread says why). */
mk_input* mk_input_open(const char* path);

/* Reads up to n decompressed bytes into dst: the number of bytes (at least 1), 0 at the end of the input, -1 on error. */
long long mk_input_read(mk_input* in, void* dst, unsigned long long n);

/* Text of the error that ended the stream ("" if none); with NULL: why the last mk_input_open of this thread failed. */
const char* mk_input_error(const mk_input* in);

void mk_input_close(mk_input* in);

#ifdef __cplusplus
}
#endif

#endif /* MERKURIO_IO_H */
