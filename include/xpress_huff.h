/*
 * xpress_huff.h -- Xpress-Huffman (MS-XCA "LZ77+Huffman", WIM, format 4) one-shot API, B200-native.
 *
 * Replaces reference include/xpress_huff.h:46-49 (implemented upstream in src/xpress_huff_compress.cpp:247-331
 * and src/xpress_huff_decompress.cpp:136-162).  xpress_huff_compress writes the very bytes of the reference's encoder
 * (its LZ77 tokens and the code lengths of HuffmanEncoder::CreateCodes, replayed tie for tie); streams of either library
 * decode with the other.
 */
#ifndef XPRESS_HUFF_H
#define XPRESS_HUFF_H
#include "mscomp/general.h"
EXTERN_C_START
MSCOMPAPI MSCompStatus xpress_huff_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI MSCompStatus xpress_huff_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI size_t xpress_huff_max_compressed_size(size_t in_len);   /* reference src/xpress_huff_compress.cpp:46 */
EXTERN_C_END
#endif
