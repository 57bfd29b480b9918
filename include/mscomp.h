/*
 * mscomp.h -- format-dispatching one-shot API of the B200-native libMSCompression.
 *
 * Replaces the reference's include/mscomp.h one-shot entry points
 *   ms_compress            (reference include/mscomp.h:59,  src/mscomp.cpp:113-117)
 *   ms_decompress          (reference include/mscomp.h:88,  src/mscomp.cpp:128-132)
 *   ms_max_compressed_size (reference include/mscomp.h:99,  src/mscomp.cpp:96-102)
 * Same names, argument order and status conventions: *out_len is the capacity of `out` on entry and the number
 * of bytes produced on success; it is unspecified on failure.  Buffers are HOST memory owned by the caller;
 * the call stages them through the GPU (see mscomp_gpu.h for the device-resident batch interface).
 * Every call is re-entrant; each calling thread gets its own CUDA stream and scratch.
 * There is no CPU implementation behind these functions: without a usable CUDA device they return
 * MSCOMP_MEM_ERROR after printing one diagnostic line to stderr.
 */
#ifndef MSCOMPRESSION_H
#define MSCOMPRESSION_H
#include "mscomp/general.h"

EXTERN_C_START

MSCOMPAPI MSCompStatus ms_compress(MSCompFormat format, const_bytes in, size_t in_len, bytes out, size_t* out_len);
MSCOMPAPI MSCompStatus ms_decompress(MSCompFormat format, const_bytes in, size_t in_len, bytes out, size_t* out_len);
/* Upper bound of ms_compress output for in_len input bytes; (size_t)-1 for an invalid format. */
MSCOMPAPI size_t ms_max_compressed_size(MSCompFormat format, size_t in_len);

/* Streaming dispatch (reference include/mscomp.h:114-213, src/mscomp.cpp:134-223): MSCOMP_NONE (copy), MSCOMP_LZNT1 and
 * MSCOMP_XPRESS -- the latter exactly as far as upstream goes: ms_inflate* work (xpress.h), ms_deflate_init answers
 * MSCOMP_MEM_ERROR because upstream never finished xpress_deflate (src/xpress_compress.cpp:49-73).  Xpress-Huffman has no
 * streaming API upstream (MSCOMP_ARG_ERROR, as for any other format number). */
MSCOMPAPI MSCompStatus ms_deflate_init(MSCompFormat format, mscomp_stream* stream);
MSCOMPAPI MSCompStatus ms_deflate(mscomp_stream* stream, MSCompFlush flush);
MSCOMPAPI MSCompStatus ms_deflate_end(mscomp_stream* stream);
MSCOMPAPI MSCompStatus ms_inflate_init(MSCompFormat format, mscomp_stream* stream);
MSCOMPAPI MSCompStatus ms_inflate(mscomp_stream* stream);
MSCOMPAPI MSCompStatus ms_inflate_end(mscomp_stream* stream);

EXTERN_C_END
#endif
