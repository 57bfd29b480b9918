/*
 * xpress.h -- plain Xpress (LZ77, MS-XCA "LZ77" variant, format 3) one-shot API, B200-native.
 *
 * Replaces reference include/xpress.h:47-50 (implemented upstream in src/xpress_compress.cpp:240-347 and
 * src/xpress_decompress.cpp:405-462).  xpress_compress writes the very bytes of the reference's encoder (its hash-chain
 * parse is rebuilt position for position, csrc/kernels/lz_exact.cuh); streams of either library decode with the other.
 */
#ifndef XPRESS_H
#define XPRESS_H
#include "mscomp/general.h"
EXTERN_C_START
MSCOMPAPI MSCompStatus xpress_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI MSCompStatus xpress_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI size_t xpress_max_compressed_size(size_t in_len);    /* reference src/xpress_compress.cpp:30 */

/* Streaming (reference include/xpress.h:52-58).
 * xpress_inflate*: src/xpress_decompress.cpp:45-61, :247-403 -- a call decodes what it is given and keeps a cut flag word /
 * symbol / match for the next call; consumed / produced / returned values equal the reference's call by call.  The stream is
 * byte-serial with an 8 KiB history, so one GPU thread walks it (state in device memory): API completeness, not a fast path.
 * xpress_deflate*: upstream never finished them; like its release build, init answers MSCOMP_MEM_ERROR, deflate
 * MSCOMP_ARG_ERROR, and deflate_end only checks its argument (src/xpress_compress.cpp:50-74, :75-221, :223-236). */
MSCOMPAPI MSCompStatus xpress_deflate_init(mscomp_stream* stream);
MSCOMPAPI MSCompStatus xpress_deflate(mscomp_stream* stream, MSCompFlush flush);
MSCOMPAPI MSCompStatus xpress_deflate_end(mscomp_stream* stream);
MSCOMPAPI MSCompStatus xpress_inflate_init(mscomp_stream* stream);
MSCOMPAPI MSCompStatus xpress_inflate(mscomp_stream* stream);
MSCOMPAPI MSCompStatus xpress_inflate_end(mscomp_stream* stream);
EXTERN_C_END
#endif
