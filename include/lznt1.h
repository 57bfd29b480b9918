/*
 * lznt1.h -- LZNT1 one-shot API (NTFS / RtlCompressBuffer format 2), B200-native.
 *
 * Replaces reference include/lznt1.h:49-52 (implemented upstream in src/lznt1_compress.cpp:233-273 and
 * src/lznt1_decompress.cpp:337).  Independent 4 KiB chunks, each decoded/encoded by its own warp / thread block.
 * lznt1_compress output is byte-identical to the reference's.
 */
#ifndef LZNT1_H
#define LZNT1_H
#include "mscomp/general.h"
EXTERN_C_START
MSCOMPAPI MSCompStatus lznt1_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI MSCompStatus lznt1_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI size_t lznt1_max_compressed_size(size_t in_len);     /* reference src/lznt1_compress.cpp:27 */

/* Streaming API (reference include/lznt1.h:55-61; src/lznt1_compress.cpp:136-231, src/lznt1_decompress.cpp:122-273): same state
 * machine -- what every call consumes, produces and returns -- with the chunk codecs running on the GPU, all the chunks a call
 * can process in one batch.  deflate output is byte-identical to the reference's. */
MSCOMPAPI MSCompStatus lznt1_deflate_init(mscomp_stream* stream);
MSCOMPAPI MSCompStatus lznt1_deflate(mscomp_stream* stream, MSCompFlush flush);
MSCOMPAPI MSCompStatus lznt1_deflate_end(mscomp_stream* stream);
MSCOMPAPI MSCompStatus lznt1_inflate_init(mscomp_stream* stream);
MSCOMPAPI MSCompStatus lznt1_inflate(mscomp_stream* stream);
MSCOMPAPI MSCompStatus lznt1_inflate_end(mscomp_stream* stream);
EXTERN_C_END
#endif
