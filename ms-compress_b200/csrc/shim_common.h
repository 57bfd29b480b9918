// shim_common.h -- declarations shared by the host-side translation units of libMSCompression.
#pragma once
#include "../../include/mscomp.h"
#include "../../include/lznt1.h"
#include "../../include/xpress.h"
#include "../../include/xpress_huff.h"
#include "../../include/mscomp_gpu.h"

// The calling thread's lazily created context (its own stream + scratch); nullptr when no CUDA device is usable.
mscomp_gpu_ctx* msc_thread_ctx();

// dst = concatenation of the first d_sizes[u] bytes of every slot; d_offs receives n+1 offsets (device).
MSCompStatus msc_compact_slots(mscomp_gpu_ctx* ctx, const uint8_t* d_slots, uint64_t stride, const uint32_t* d_sizes,
                               uint32_t n, uint64_t* d_offs, uint8_t* d_dst);

// LZNT1 encoder over a slice of chunks (lets the host wrapper overlap uploads with encoding)
MSCompStatus msc_lznt1_encode_chunks(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t first, uint32_t count,
                                     uint8_t* d_slots, uint32_t* d_sizes);
constexpr size_t MSC_LZNT1_SLOT_BYTES = 4112;

// reduce per-unit sizes / statuses into one mscomp_gpu_result (device)
MSCompStatus msc_summarize(mscomp_gpu_ctx* ctx, const uint32_t* d_sizes, const int32_t* d_status, uint32_t n, mscomp_gpu_result* d_res);

// Xpress-Huffman encoder over explicit units; eos_mode 1 = the units are consecutive chunks of one stream.
extern "C" MSCompStatus msc_xpress_huff_compress_units(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b, int eos_mode);

#ifdef MSC_NO_XPRESS   // bring-up only
#define MSC_XPRESS_COMPRESS nullptr
#define MSC_XPRESS_DECOMPRESS nullptr
#define MSC_XPRESS_MAX nullptr
#define MSC_XPRESS_HUFF_COMPRESS nullptr
#define MSC_XPRESS_HUFF_DECOMPRESS nullptr
#define MSC_XPRESS_HUFF_MAX nullptr
#else
#define MSC_XPRESS_COMPRESS xpress_compress
#define MSC_XPRESS_DECOMPRESS xpress_decompress
#define MSC_XPRESS_MAX xpress_max_compressed_size
#define MSC_XPRESS_HUFF_COMPRESS xpress_huff_compress
#define MSC_XPRESS_HUFF_DECOMPRESS xpress_huff_decompress
#define MSC_XPRESS_HUFF_MAX xpress_huff_max_compressed_size
#endif
