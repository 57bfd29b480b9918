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

MSCompStatus msc_compact_slots_range(mscomp_gpu_ctx* ctx, const uint8_t* d_slots, uint64_t stride, const uint32_t* d_sizes,
                                     uint32_t first, uint32_t n, uint64_t* d_offs, uint8_t* d_dst, uint64_t* h_end);

// LZNT1 encoder over a slice of chunks (lets the host wrapper overlap uploads with encoding)
MSCompStatus msc_lznt1_encode_chunks(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t first, uint32_t count,
                                     uint8_t* d_slots, uint32_t* d_sizes);
constexpr size_t MSC_LZNT1_SLOT_BYTES = 4112;

// LZNT1 chunk index built on the GPU (kernels/lznt1_index.cuh), range by range; layout == msc::Lznt1IndexState
struct msc_lznt1_index_state { uint32_t entry, done, tail, next_entry; uint64_t ip, nchunks, range_chunks; };
constexpr size_t MSC_LZNT1_INDEX_SEG = 65536;          // segment size of the index builder (ranges are multiples of it)
// nseg_cap = the largest range of the pipeline (fixes the layout of the two scratch slots)
extern "C" MSCompStatus msc_lznt1_index_count(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t seg0, uint32_t nseg, uint32_t nseg_cap,
                                              int slot, msc_lznt1_index_state* d_state, msc_lznt1_index_state* h_mirror);
extern "C" MSCompStatus msc_lznt1_index_write(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t seg0, uint32_t nseg, uint32_t nseg_cap,
                                              int slot, uint64_t* d_offs);

extern "C" MSCompStatus msc_lznt1_decode_chunks(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, const uint64_t* d_chunk_off, uint32_t nchunks,
                                                uint8_t* d_out, uint32_t* sizes, int32_t* status);
extern "C" MSCompStatus msc_summarize_units(mscomp_gpu_ctx* ctx, const uint32_t* d_sizes, const int32_t* d_status, uint32_t n, uint32_t full_unit, mscomp_gpu_result* d_res);

// reduce per-unit sizes / statuses into one mscomp_gpu_result (device)
MSCompStatus msc_summarize(mscomp_gpu_ctx* ctx, const uint32_t* d_sizes, const int32_t* d_status, uint32_t n, mscomp_gpu_result* d_res);

// Xpress-Huffman encoder over explicit units; eos_mode 1 = the units are consecutive chunks of one stream.
extern "C" MSCompStatus msc_xpress_huff_compress_units(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b, int eos_mode);

// One multi-chunk Xpress-Huffman stream on the device: chunk-parallel parse + token execution (kernels/xhuff_chunks.cuh);
// *handled = false leaves the stream to the serial decoder.
extern "C" MSCompStatus msc_xhuff_stream_decode(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint32_t in_len, uint8_t* d_out, uint64_t cap, uint64_t* total, bool* handled);
extern "C" MSCompStatus msc_xpress_stream_decode(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint32_t in_len, uint8_t* d_out, uint64_t cap, uint64_t* total, bool* handled);

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
