/*
 * xpress.h -- plain Xpress (LZ77, MS-XCA "LZ77" variant, format 3) one-shot API, B200-native.
 *
 * Replaces reference include/xpress.h:47-50 (implemented upstream in src/xpress_compress.cpp:240-347 and
 * src/xpress_decompress.cpp:405-462).  Our compressed bytes differ from the reference's where match choice
 * differs; they decode with the reference's xpress_decompress and vice versa.
 */
#ifndef XPRESS_H
#define XPRESS_H
#include "mscomp/general.h"
EXTERN_C_START
MSCOMPAPI MSCompStatus xpress_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI MSCompStatus xpress_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI size_t xpress_max_compressed_size(size_t in_len);    /* reference src/xpress_compress.cpp:30 */
EXTERN_C_END
#endif
