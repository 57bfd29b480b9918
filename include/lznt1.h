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
EXTERN_C_END
#endif
