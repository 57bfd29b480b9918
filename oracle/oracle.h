/*
 * oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's (coderforlife/ms-compress) LZNT1 / Xpress / Xpress-Huffman
 * one-shot codecs, written from the behaviour of the files cited at each function.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the
 * product (ms-compress_b200/) never does.
 *
 * Parity pinning: the reference ships no vectors, so this oracle is pinned by
 *   (1) the five MS-XCA documentation vectors (tests/golden/kat.json),
 *   (2) byte-for-byte agreement with the compiled reference (oracle/_ref/libMSCompression.so) on
 *       seeded generator data in both directions (tests/test_oracle.py, runs wherever _ref exists),
 *   (3) committed golden fixtures produced by the compiled reference (tests/golden/, made by
 *       tests/golden/make_golden.py).
 *
 * Status codes mirror include/mscomp/general.h:74-85 of the reference.
 */
#ifndef MSCOMP_ORACLE_H
#define MSCOMP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_OK = 0,
    ORC_ARG_ERROR = -2,
    ORC_DATA_ERROR = -3,
    ORC_MEM_ERROR = -4,
    ORC_BUF_ERROR = -5
};

/* LZNT1 -- src/lznt1_compress.cpp, src/lznt1_decompress.cpp, include/mscomp/LZNT1Dictionary.h */
size_t orc_lznt1_max_compressed_size(size_t in_len);
int orc_lznt1_compress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len);
int orc_lznt1_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len);
/* one chunk payload (after the 2-byte header) -> up to 4096 bytes; src/lznt1_decompress.cpp:37-121 */
int orc_lznt1_decompress_chunk(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len);

/* Xpress -- src/xpress_compress.cpp, src/xpress_decompress.cpp, include/mscomp/XpressDictionary.h */
size_t orc_xpress_max_compressed_size(size_t in_len);
int orc_xpress_compress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len);
int orc_xpress_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len);

/* Xpress-Huffman -- src/xpress_huff_compress.cpp, src/xpress_huff_decompress.cpp,
 * include/mscomp/{HuffmanEncoder,HuffmanDecoder,Bitstream}.h */
size_t orc_xpress_huff_max_compressed_size(size_t in_len);
int orc_xpress_huff_compress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len);
int orc_xpress_huff_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len);

#ifdef __cplusplus
}
#endif
#endif
