// mscomp_gpu.cu -- device-resident batch API (include/mscomp_gpu.h): context management and kernel launches.
#include "ctx.cuh"
#include "shim_common.h"
#include "kernels/lznt1_decode.cuh"
#include "kernels/lznt1_encode.cuh"
#include "kernels/lznt1_index.cuh"
#include <stdlib.h>
#include <string.h>
#include "kernels/compact.cuh"
#include "kernels/summary.cuh"
#include "kernels/xhuff_decode.cuh"
#include "kernels/xhuff_index.cuh"
#include "kernels/xhuff_decode_indexed.cuh"
#include "kernels/xhuff_chunks.cuh"
#include "kernels/xhuff_chunk_warp.cuh"
#include "kernels/lz_jump.cuh"
#include "kernels/xpress_decode.cuh"
#include "kernels/xhuff_encode.cuh"
#include "kernels/xpress_encode.cuh"
#include "kernels/xpress_spec.cuh"
#include <new>
#include <algorithm>
#include <vector>

using namespace msc;

static_assert(sizeof(GpuResult) == sizeof(mscomp_gpu_result), "result layout");

// Units [first, first + n) of a slot array: offsets continue from d_offs[first] (written by the previous range; 0 when
// first == 0), the bytes go to their final place in d_dst.  A whole array is the range (0, n).
MSCompStatus msc_compact_slots_range(mscomp_gpu_ctx* ctx, const uint8_t* d_slots, uint64_t stride, const uint32_t* d_sizes,
                                     uint32_t first, uint32_t n, uint64_t* d_offs, uint8_t* d_dst, uint64_t* h_end)
{
    MSC_LAUNCH(ctx, scan_sizes_kernel, 1, SCAN_THREADS, d_sizes + first, n, (uint64_t)0, first ? (const uint64_t*)(d_offs + first) : nullptr, d_offs + first, h_end);
    if (n) {
        const uint32_t gw = (n + 7) / 8;
        const uint32_t grid = gw < (uint32_t)ctx->sm_count * 8 ? gw : (uint32_t)ctx->sm_count * 8;
        MSC_LAUNCH(ctx, gather_slots_kernel, grid, 256, d_slots + (uint64_t)first * stride, stride, d_sizes + first, (const uint64_t*)(d_offs + first), n, d_dst);
    }
    return msc_cuda_status(cudaGetLastError(), "compact_slots");
}

MSCompStatus msc_compact_slots(mscomp_gpu_ctx* ctx, const uint8_t* d_slots, uint64_t stride, const uint32_t* d_sizes,
                               uint32_t n, uint64_t* d_offs, uint8_t* d_dst)
{
    return msc_compact_slots_range(ctx, d_slots, stride, d_sizes, 0, n, d_offs, d_dst, nullptr);
}

// Encode chunks [first, first+count) of the stream d_in[0..in_len) into their slots (one 256-thread block per chunk).
MSCompStatus msc_lznt1_encode_chunks(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t first, uint32_t count,
                                     uint8_t* d_slots, uint32_t* d_sizes)
{
    // one block per chunk: the hardware block scheduler balances the chunks (their cost varies with the data) better than a
    // static grid-stride assignment; a block lives ~50 us, so its launch cost does not matter
    const uint32_t grid = count;
    MSC_OPT_IN_SMEM(lznt1_encode_kernel, sizeof(Lznt1EncSmem));
    static const char* hw_env = getenv("MSCOMP_LZNT1_HEAVY_WORK");         // tuning / test hook: candidate evaluations above which a chunk is "heavy"
    const uint32_t heavy_work = hw_env ? (uint32_t)strtoul(hw_env, nullptr, 10) : LZE_HEAVY_WORK;
    MSC_LAUNCH_SMEM(ctx, lznt1_encode_kernel, grid, LZE_THREADS, sizeof(Lznt1EncSmem), d_in, (uint64_t)in_len, first, count, d_slots, d_sizes, heavy_work);
    return msc_cuda_status(cudaGetLastError(), "lznt1_encode");
}

MSCompStatus msc_summarize(mscomp_gpu_ctx* ctx, const uint32_t* d_sizes, const int32_t* d_status, uint32_t n, mscomp_gpu_result* d_res)
{
    MSC_LAUNCH(ctx, summarize_kernel, 1, SUM_THREADS, d_sizes, d_status, n, 0u, (GpuResult*)d_res);
    return msc_cuda_status(cudaGetLastError(), "summarize");
}

extern "C" {

static mscomp_gpu_ctx* create_ctx(int device, void* cuda_stream, bool use_given)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    mscomp_gpu_ctx* ctx = new (std::nothrow) mscomp_gpu_ctx();
    if (!ctx) return nullptr;
    ctx->device = device;
    if (use_given) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete ctx; return nullptr; }
        ctx->own_stream = true;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (ctx->sm_count <= 0) ctx->sm_count = 148;
    const char* enc = getenv("MSCOMP_GPU_ENCODER");
    ctx->encoder = (enc && strcmp(enc, "fast") == 0) ? MSCOMP_GPU_ENCODER_FAST : MSCOMP_GPU_ENCODER_REFERENCE;
    return ctx;
}

mscomp_gpu_ctx* mscomp_gpu_create(int device) { return create_ctx(device, nullptr, false); }
mscomp_gpu_ctx* mscomp_gpu_create_on_stream(int device, void* cuda_stream) { return create_ctx(device, cuda_stream, true); }

void mscomp_gpu_destroy(mscomp_gpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    mscomp_gpu_comm_release(ctx);
    for (int i = 0; i < SC_COUNT; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_big) cudaFreeHost(ctx->pinned_big);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->copy_stream) {
        for (int i = 0; i < 16; ++i) { cudaEventDestroy(ctx->copy_ev[i]); cudaEventDestroy(ctx->work_ev[i]); }
        cudaStreamDestroy(ctx->copy_stream); cudaStreamDestroy(ctx->down_stream);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

MSCompStatus mscomp_gpu_synchronize(mscomp_gpu_ctx* ctx)
{
    if (!ctx) return MSCOMP_ARG_ERROR;
    return msc_cuda_status(cudaStreamSynchronize(ctx->stream), "synchronize");
}

uint64_t mscomp_gpu_kernel_launches(const mscomp_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

void mscomp_gpu_set_encoder(mscomp_gpu_ctx* ctx, int mode) { if (ctx) ctx->encoder = mode == MSCOMP_GPU_ENCODER_FAST ? MSCOMP_GPU_ENCODER_FAST : MSCOMP_GPU_ENCODER_REFERENCE; }
int mscomp_gpu_get_encoder(const mscomp_gpu_ctx* ctx) { return ctx ? ctx->encoder : 0; }

void mscomp_gpu_profile_enable(mscomp_gpu_ctx* ctx, int on)
{
    if (!ctx) return;
    ctx->prof_on = on != 0;
    ctx->prof_n = 0;
}

// Synchronises, then writes "kernel_name calls total_ms" lines (one per distinct kernel) into buf; returns bytes written.
size_t mscomp_gpu_profile_report(mscomp_gpu_ctx* ctx, char* buf, size_t cap)
{
    if (!ctx || !buf || !cap) return 0;
    cudaStreamSynchronize(ctx->stream);
    size_t w = 0;
    buf[0] = 0;
    for (int i = 0; i < ctx->prof_n; ++i) {
        bool seen = false;
        for (int j = 0; j < i; ++j) if (ctx->prof_name[j] == ctx->prof_name[i]) { seen = true; break; }
        if (seen) continue;
        double ms = 0; int calls = 0;
        for (int j = i; j < ctx->prof_n; ++j) if (ctx->prof_name[j] == ctx->prof_name[i]) {
            float t = 0; cudaEventElapsedTime(&t, ctx->prof_ev[2 * j], ctx->prof_ev[2 * j + 1]); ms += t; ++calls;
        }
        int k = snprintf(buf + w, cap - w, "%s %d %.6f\n", ctx->prof_name[i], calls, ms);
        if (k < 0 || (size_t)k >= cap - w) break;
        w += (size_t)k;
    }
    ctx->prof_n = 0;
    return w;
}

// ---- LZNT1 ------------------------------------------------------------------------------------------------------

MSCompStatus mscomp_gpu_lznt1_compress(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                                       uint64_t* d_chunk_off, mscomp_gpu_result* d_result)
{
    if (!ctx || (!d_in && in_len) || !d_out || !d_result) return MSCOMP_ARG_ERROR;
    const uint64_t nchunks64 = (in_len + LZNT1_CHUNK - 1) / LZNT1_CHUNK;
    if (nchunks64 > 0x7fffffffu) return MSCOMP_ARG_ERROR;
    const uint32_t nchunks = (uint32_t)nchunks64;
    if (out_cap < in_len + 2 * nchunks64) return MSCOMP_BUF_ERROR;            // worst case: every chunk stored raw
    cudaSetDevice(ctx->device);
    uint8_t* slots = (uint8_t*)ctx->get(SC_SLOTS, (size_t)nchunks * LZNT1_SLOT + 16);
    uint32_t* sizes = (uint32_t*)ctx->get(SC_SIZES, ((size_t)nchunks + 1) * sizeof(uint32_t));
    uint64_t* offs = d_chunk_off ? d_chunk_off : (uint64_t*)ctx->get(SC_OFFS, ((size_t)nchunks + 1) * sizeof(uint64_t));
    if (!slots || !sizes || !offs) return MSCOMP_MEM_ERROR;
    if (nchunks) msc_lznt1_encode_chunks(ctx, d_in, in_len, 0, nchunks, slots, sizes);
    msc_compact_slots(ctx, slots, LZNT1_SLOT, sizes, nchunks, offs, d_out);
    // result: bytes = offs[nchunks]
    cudaMemsetAsync(d_result, 0, sizeof(mscomp_gpu_result), ctx->stream);
    cudaMemcpyAsync(&d_result->bytes, offs + nchunks, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream);
    return msc_cuda_status(cudaGetLastError(), "lznt1_compress");
}

// ---- LZNT1 chunk index (kernels/lznt1_index.cuh) -------------------------------------------------------------------
static_assert(sizeof(Lznt1IndexState) == sizeof(msc_lznt1_index_state), "index state layout");

// Steps 1-4 for segments [seg0, seg0 + nseg) of the stream: d_state is updated (chunks of the range, how the chain
// ended).  The per-segment entries / chunk numbers stay in scratch slot `slot` (0..7) for the matching _write call;
// h_mirror (may be null): mapped host memory that receives a copy of the state.
constexpr int LZI_SLOTS = 8;          // ranges whose count may be pending between _count and _write
// Bytes of one scratch slot (seg_base u32 | seg_entry u16 | grp_entry u16) for ranges of up to nseg_cap segments.
static size_t lzi_slot_bytes(uint32_t nseg_cap)
{
    const size_t ng = (nseg_cap + LZI_GROUP - 1) / LZI_GROUP;
    return (((size_t)nseg_cap * 6 + ng * 2) + 255) / 256 * 256;
}

MSCompStatus msc_lznt1_index_count(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t seg0, uint32_t nseg, uint32_t nseg_cap,
                                   int slot, msc_lznt1_index_state* d_state, msc_lznt1_index_state* h_mirror)
{
    if (nseg > nseg_cap) return MSCOMP_ARG_ERROR;
    const uint32_t S = LZI_SEG_BYTES, ng = (nseg + LZI_GROUP - 1) / LZI_GROUP;
    const size_t t_bytes = ((size_t)nseg + ng) * LZI_STRIDE * sizeof(uint16_t);
    const size_t e_bytes = lzi_slot_bytes(nseg_cap);                                       // the slot layout must not depend on this range's size
    uint16_t* T1 = (uint16_t*)ctx->get(SC_AUX0, t_bytes + 64);
    uint8_t* e = (uint8_t*)ctx->get(SC_AUX1, LZI_SLOTS * e_bytes + 64);
    if (!T1 || !e) return MSCOMP_MEM_ERROR;
    uint16_t* T2 = T1 + (size_t)nseg * LZI_STRIDE;
    e += (size_t)(slot % LZI_SLOTS) * e_bytes;
    uint32_t* seg_base = (uint32_t*)e;
    uint16_t* seg_entry = (uint16_t*)(seg_base + nseg);
    uint16_t* grp_entry = seg_entry + nseg;
    const uint32_t grid = nseg < (uint32_t)ctx->sm_count * 12 ? nseg : (uint32_t)ctx->sm_count * 12;
    MSC_OPT_IN_SMEM(lznt1_index_tables_kernel, S + 16);
    MSC_LAUNCH_SMEM(ctx, lznt1_index_tables_kernel, grid, LZI_THREADS, S + 16, d_in, (uint64_t)in_len, seg0, nseg, S, T1);
    MSC_LAUNCH(ctx, lznt1_index_compose_kernel, ng, LZI_THREADS, (const uint16_t*)T1, nseg, T2);
    MSC_LAUNCH(ctx, lznt1_index_entries_kernel, 1, 256, (const uint16_t*)T1, (const uint16_t*)T2, nseg, (Lznt1IndexState*)d_state, seg_entry, grp_entry);
    MSC_LAUNCH(ctx, lznt1_index_count_kernel, 1, LZI_THREADS, d_in, (uint64_t)in_len, seg0, nseg, S, (const uint16_t*)seg_entry, seg_base, (Lznt1IndexState*)d_state, (Lznt1IndexState*)h_mirror);
    return msc_cuda_status(cudaGetLastError(), "lznt1_index");
}

// Step 5: chunk offsets of the range into d_offs[0 .. range_chunks) (the caller passes the position of the range's first chunk).
MSCompStatus msc_lznt1_index_write(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint32_t seg0, uint32_t nseg, uint32_t nseg_cap,
                                   int slot, uint64_t* d_offs)
{
    const uint32_t S = LZI_SEG_BYTES;
    const size_t e_bytes = lzi_slot_bytes(nseg_cap);
    uint8_t* e = (uint8_t*)ctx->scratch[SC_AUX1];
    if (!e) return MSCOMP_ARG_ERROR;
    e += (size_t)(slot % LZI_SLOTS) * e_bytes;
    const uint32_t* seg_base = (const uint32_t*)e;
    const uint16_t* seg_entry = (const uint16_t*)(seg_base + nseg);
    const uint32_t grid = (nseg + 255) / 256;
    MSC_LAUNCH(ctx, lznt1_index_write_kernel, grid, 256, d_in, (uint64_t)in_len, seg0, nseg, S, seg_entry, seg_base, d_offs);
    return msc_cuda_status(cudaGetLastError(), "lznt1_index");
}

MSCompStatus mscomp_gpu_lznt1_index(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint64_t* d_chunk_off, size_t chunk_cap,
                                    mscomp_gpu_lznt1_index_result* h_result)
{
    if (!ctx || (!d_in && in_len) || !h_result || (!d_chunk_off && chunk_cap)) return MSCOMP_ARG_ERROR;
    cudaSetDevice(ctx->device);
    const uint64_t nseg64 = in_len ? (in_len + LZI_SEG_BYTES - 1) / LZI_SEG_BYTES : 1;
    if (nseg64 > 0x7fffffffu) return MSCOMP_ARG_ERROR;
    msc_lznt1_index_state* d_st = (msc_lznt1_index_state*)ctx->get(SC_RESULT, 256);
    msc_lznt1_index_state* h_st = (msc_lznt1_index_state*)ctx->get_pinned(256);
    if (!d_st || !h_st) return MSCOMP_MEM_ERROR;
    cudaMemsetAsync(d_st, 0, sizeof(*d_st), ctx->stream);
    MSCompStatus st = msc_lznt1_index_count(ctx, d_in, in_len, 0, (uint32_t)nseg64, (uint32_t)nseg64, 0, d_st, nullptr);
    if (st != MSCOMP_OK) return st;
    cudaMemcpyAsync(h_st, d_st, sizeof(*h_st), cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_index")) != MSCOMP_OK) return st;
    h_result->nchunks = h_st->nchunks;
    h_result->consumed = h_st->ip;
    h_result->tail = (int32_t)h_st->tail;
    h_result->reserved = 0;
    if (h_st->nchunks > chunk_cap) return MSCOMP_BUF_ERROR;
    if (h_st->nchunks) {
        if ((st = msc_lznt1_index_write(ctx, d_in, in_len, 0, (uint32_t)nseg64, (uint32_t)nseg64, 0, d_chunk_off)) != MSCOMP_OK) return st;
        st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_index");
    }
    return st;
}

// Decode chunks whose headers sit at d_chunk_off[0 .. nchunks) of the stream; chunk i goes to d_out + 4096 * i.
MSCompStatus msc_lznt1_decode_chunks(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, const uint64_t* d_chunk_off, uint32_t nchunks,
                                     uint8_t* d_out, uint32_t* sizes, int32_t* status)
{
    // A warp per chunk, the chunk decoded in place in shared memory (6 KB per warp): DRAM traffic equals the algorithmic
    // bytes.  Persistent blocks: 8 per SM (32 warps, the register budget of the kernel), each warp strides over the chunks.
    static int per_sm = 0;
    if (!per_sm) {
        cudaFuncSetAttribute(lznt1_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lznt1_decode_kernel, LZNT1_WARPS_PER_BLOCK * 32, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 8; }
    }
    const uint32_t blocks = (nchunks + LZNT1_WARPS_PER_BLOCK - 1) / LZNT1_WARPS_PER_BLOCK;
    const uint32_t cap = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
    MSC_LAUNCH(ctx, lznt1_decode_kernel, blocks < cap ? blocks : cap, LZNT1_WARPS_PER_BLOCK * 32,
               d_in, (uint64_t)in_len, d_chunk_off, nchunks, d_out, sizes, status);
    return msc_cuda_status(cudaGetLastError(), "lznt1_decode");
}

MSCompStatus msc_summarize_units(mscomp_gpu_ctx* ctx, const uint32_t* d_sizes, const int32_t* d_status, uint32_t n, uint32_t full_unit, mscomp_gpu_result* d_res)
{
    MSC_LAUNCH(ctx, summarize_kernel, 1, SUM_THREADS, d_sizes, d_status, n, full_unit, (GpuResult*)d_res);
    return msc_cuda_status(cudaGetLastError(), "summarize");
}

MSCompStatus mscomp_gpu_lznt1_decompress(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, const uint64_t* d_chunk_off,
                                         uint32_t nchunks, uint8_t* d_out, uint32_t* d_chunk_size, int32_t* d_chunk_status,
                                         mscomp_gpu_result* d_result)
{
    if (!ctx || !d_result || (nchunks && (!d_in || !d_chunk_off || !d_out))) return MSCOMP_ARG_ERROR;
    cudaSetDevice(ctx->device);
    uint32_t* sizes = d_chunk_size ? d_chunk_size : (uint32_t*)ctx->get(SC_SIZES, ((size_t)nchunks + 1) * sizeof(uint32_t));
    int32_t* status = d_chunk_status ? d_chunk_status : (int32_t*)ctx->get(SC_STATUS, ((size_t)nchunks + 1) * sizeof(int32_t));
    if (!sizes || !status) return MSCOMP_MEM_ERROR;
    if (nchunks) msc_lznt1_decode_chunks(ctx, d_in, in_len, d_chunk_off, nchunks, d_out, sizes, status);
    MSC_LAUNCH(ctx, summarize_kernel, 1, SUM_THREADS, sizes, status, nchunks, (uint32_t)LZNT1_CHUNK, (GpuResult*)d_result);
    return msc_cuda_status(cudaGetLastError(), "lznt1_decompress");
}


// ---- Xpress-Huffman / Xpress batches ---------------------------------------------------------------------------------

// Token scratch for the two-kernel decoders: 4 bytes per output byte of the region the batch can touch.  Returns the
// scratch pointer, or nullptr when the region is unknown / too large (the single-kernel decoder is used then).
static uint32_t* token_scratch(mscomp_gpu_ctx* ctx, const mscomp_gpu_batch* b)
{
    static const char* mode = getenv("MSCOMP_LZ_DECODE");            // "fused" forces the single-kernel decoders
    if (mode && strcmp(mode, "fused") == 0) return nullptr;
    uint64_t* d_ext = (uint64_t*)ctx->get(SC_RESULT, 256);
    uint64_t* h_ext = (uint64_t*)ctx->get_pinned(256);
    if (!d_ext || !h_ext) return nullptr;
    MSC_LAUNCH(ctx, batch_extent_kernel, 1, SUM_THREADS, b->d_out_off, b->d_out_cap, b->count, d_ext + 8);
    cudaMemcpyAsync(h_ext + 8, d_ext + 8, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nullptr;
    const uint64_t extent = h_ext[8];
    const uint64_t have = ctx->scratch_cap[SC_AUX2];
    if (extent == 0 || extent > (1ull << 36)) return nullptr;
    if (4 * extent + 64 > have) {                                     // growing: only if it fits comfortably
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (4 * extent + 64 > (uint64_t)free_b / 2 + have) return nullptr;
    }
    return (uint32_t*)ctx->get(SC_AUX2, 4 * extent + 64);
}

static bool batch_ok(const mscomp_gpu_ctx* ctx, const uint8_t* d_in, const uint8_t* d_out, const mscomp_gpu_batch* b)
{
    return ctx && b && (b->count == 0 || (d_in && d_out && b->d_in_off && b->d_in_len && b->d_out_off && b->d_out_cap && b->d_out_len && b->d_status));
}

static MSCompStatus xhuff_encode_launch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b, int eos_mode,
                                        XhSeg* d_index, const uint64_t* d_index_off, uint32_t* d_nseg)
{
    if (!batch_ok(ctx, d_in, d_out, b)) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    // Default: the reference's LZ77 tokens (lz_exact.cuh: hash-chain links, Find, window 65535 across the chunks of a stream) and code
    // lengths.  MSCOMP_GPU_ENCODER_FAST selects the earlier bulk-synchronous matcher (window reset per chunk; valid streams, other bytes).
    const bool exact = ctx->encoder != MSCOMP_GPU_ENCODER_FAST;
    XhPre pre{};
    if (exact) {
        uint32_t* seg_first = (uint32_t*)ctx->get(SC_AUX1, ((size_t)b->count + 2) * sizeof(uint32_t) + 128);
        unsigned long long* d_tot = (unsigned long long*)ctx->get(SC_RESULT, 1024);
        unsigned long long* h_tot = (unsigned long long*)ctx->get_pinned(1024);
        if (!seg_first || !d_tot || !h_tot) return MSCOMP_MEM_ERROR;
        d_tot += 64; h_tot += 64;                                      // (clear of the result mailbox at the front)
        uint32_t* work = (uint32_t*)(d_tot + 4);
        unsigned long long total_segs = 0;
        bool big = false;                                              // units of more than one chunk: links reach back across chunks
        uint32_t fcount = b->count;
        const uint64_t* f_off = b->d_in_off;
        const uint32_t* f_len = b->d_in_len;
        if (eos_mode == 1) {                                           // the units are the chunks of ONE stream: Find sees the stream
            uint64_t* desc_off = (uint64_t*)(d_tot + 8);
            uint32_t* desc_len = (uint32_t*)(d_tot + 9);
            MSC_LAUNCH(ctx, xhuff_stream_desc_kernel, 1, 32, b->count, b->d_in_off, b->d_in_len, desc_off, desc_len, seg_first, work);
            total_segs = b->count; big = b->count > 1; fcount = 1; f_off = desc_off; f_len = desc_len;
            pre.stream_len = desc_len;
        } else {
            MSC_LAUNCH(ctx, xpress_segments_kernel, 1, 1024, b->d_in_len, b->count, seg_first, d_tot, work);
            cudaMemcpyAsync(h_tot, d_tot, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return msc_cuda_status(cudaGetLastError(), "xpress_huff_compress");
            total_segs = h_tot[0]; big = h_tot[2] > XHE_CHUNK;
        }
        if (total_segs >= (1ull << 24)) return MSCOMP_ARG_ERROR;
        const size_t npos = (size_t)total_segs * LZX_SEG, tiles = npos / LZF_TILE;
        const size_t lo_b = npos * 4, v_b = npos / 8, x_b = tiles * 4 + (npos / LZX_SUB) * 4;    // answers, token starts, exits per tile + per sub-tile
        uint8_t* sc = (uint8_t*)ctx->get(SC_AUX4, lo_b + v_b + x_b + npos * 2 + 256);
        if (!sc) return MSCOMP_MEM_ERROR;
        uint32_t* lenoff = (uint32_t*)sc;
        uint32_t* visited0 = (uint32_t*)(sc + lo_b);
        uint32_t* exit0 = (uint32_t*)(sc + lo_b + v_b);
        uint16_t* dist = (uint16_t*)(sc + lo_b + v_b + x_b);
        pre.seg_first = seg_first; pre.lenoff = lenoff; pre.visited0 = visited0; pre.exit0 = exit0;
        if (total_segs) {
            if (big) {
                const size_t prev_smem = 32768 * sizeof(uint32_t) + LZX_STAGE + 128;
                MSC_OPT_IN_SMEM(lz_prev_kernel<uint32_t>, prev_smem);
                const uint32_t pg = total_segs < (unsigned long long)ctx->sm_count ? (uint32_t)total_segs : (uint32_t)ctx->sm_count;
                MSC_LAUNCH_SMEM(ctx, lz_prev_kernel<uint32_t>, pg, 32, prev_smem, d_in, fcount, f_off, f_len, (const uint32_t*)seg_first, work,
                                dist, 65535u, LZX_SEG);
            } else {
                const size_t prev_smem = 32768 * sizeof(uint16_t) + LZX_STAGE + 128;
                MSC_OPT_IN_SMEM(lz_prev_kernel<uint16_t>, prev_smem);
                const uint32_t pg = total_segs < (unsigned long long)ctx->sm_count * 3 ? (uint32_t)total_segs : (uint32_t)ctx->sm_count * 3;
                MSC_LAUNCH_SMEM(ctx, lz_prev_kernel<uint16_t>, pg, 32, prev_smem, d_in, fcount, f_off, f_len, (const uint32_t*)seg_first, work,
                                dist, 65535u, LZX_SEG);
            }
            cudaMemsetAsync(lenoff, 0xFF, lo_b, ctx->stream);           // LZX_UNKNOWN: nobody has asked
            const uint32_t fg = total_segs < (unsigned long long)ctx->sm_count * LZX_FIND_MINB ? (uint32_t)total_segs : (uint32_t)ctx->sm_count * LZX_FIND_MINB;
            MSC_LAUNCH(ctx, lz_find_kernel<LZX_FIND_MINB>, fg, LZX_FIND_THREADS, d_in, fcount, f_off, f_len, (const uint32_t*)seg_first,
                       work + 1, (const uint16_t*)dist, lenoff, visited0, exit0 + tiles, 65535u, XHE_CHUNK);
            const unsigned long long sb = (tiles + 127) / 128;
            const uint32_t qg = sb < (unsigned long long)ctx->sm_count * 16 ? (uint32_t)sb : (uint32_t)ctx->sm_count * 16;
            MSC_LAUNCH(ctx, lz_stitch_kernel, qg, 128, d_in, fcount, f_off, f_len, (const uint32_t*)seg_first, (const uint16_t*)dist, lenoff,
                       visited0, (const uint32_t*)(exit0 + tiles), exit0, 65535u, XHE_CHUNK);
        }
        pre.dist = dist;
    }
    MSC_OPT_IN_SMEM(xhuff_encode_kernel, sizeof(XhEncSmem));
    MSC_OPT_IN_SMEM(xhuff_encode_exact_kernel, sizeof(XhEncSmem));
    static int per_sm_of[2] = {0, 0};                                 // resident blocks per SM of the kernel that runs (per-block scratch: grid = ONE wave;
    int& per_sm = per_sm_of[exact ? 1 : 0];                           // the two kernels differ: 4 and 5 -- a grid sized for the other one leaves a tail wave)
    if (!per_sm) {
        const cudaError_t oe = exact ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xhuff_encode_exact_kernel, XHE_THREADS, sizeof(XhEncSmem))
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xhuff_encode_kernel, XHE_THREADS, sizeof(XhEncSmem));
        if (oe != cudaSuccess || per_sm < 1) per_sm = 3;
    }
    const uint32_t cap = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
    const uint32_t grid = b->count < cap ? b->count : cap;
    uint8_t* scratch = (uint8_t*)ctx->get(SC_AUX2, (size_t)grid * XHE_SCRATCH_BYTES + 64);
    if (!scratch) return MSCOMP_MEM_ERROR;
    if (exact)
        MSC_LAUNCH_SMEM(ctx, xhuff_encode_exact_kernel, grid, XHE_THREADS, sizeof(XhEncSmem), d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off,
                        b->d_out_cap, b->d_out_len, b->d_status, scratch, eos_mode, d_index, d_index_off, d_nseg, pre);
    else
        MSC_LAUNCH_SMEM(ctx, xhuff_encode_kernel, grid, XHE_THREADS, sizeof(XhEncSmem), d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off,
                        b->d_out_cap, b->d_out_len, b->d_status, scratch, eos_mode, d_index, d_index_off, d_nseg);
    return msc_cuda_status(cudaGetLastError(), "xpress_huff_compress");
}

MSCompStatus msc_xpress_huff_compress_units(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b, int eos_mode)
{
    return xhuff_encode_launch(ctx, d_in, d_out, b, eos_mode, nullptr, nullptr, nullptr);
}

MSCompStatus mscomp_gpu_xpress_huff_compress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b)
{
    return xhuff_encode_launch(ctx, d_in, d_out, b, 0, nullptr, nullptr, nullptr);
}

// The same streams, byte for byte, plus the side-band restart index of each (the encoder knows every symbol's bit address, so the
// index costs it nothing; what the serial walk of mscomp_gpu_xpress_huff_index_batch would find, entry for entry).
MSCompStatus mscomp_gpu_xpress_huff_compress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b,
                                                           mscomp_gpu_xhuff_seg* d_index, const uint64_t* d_index_off, uint32_t* d_nseg)
{
    if (b && b->count && (!d_index || !d_index_off || !d_nseg)) return MSCOMP_ARG_ERROR;
    return xhuff_encode_launch(ctx, d_in, d_out, b, 0, (XhSeg*)d_index, d_index_off, d_nseg);
}

// ---- Xpress-Huffman decode: restart index + one lane per segment ---------------------------------------------------------
static_assert(sizeof(XhSeg) == sizeof(mscomp_gpu_xhuff_seg), "index entry layout");

namespace msc {
// room (entries) the index of every stream can need: bounded by its output capacity and by its input bits
__global__ void xhi_room_kernel(const uint32_t* __restrict__ in_len, const uint32_t* __restrict__ out_cap, uint32_t n, uint32_t* __restrict__ room)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t by_out = xhi_max_segments(out_cap[i]);
    const uint64_t by_in = (uint64_t)in_len[i] / 8 + (uint64_t)in_len[i] / 260 + 3;      // >= 1 bit per symbol, >= 260 bytes per chunk
    const uint64_t r = by_out < by_in ? by_out : by_in;
    room[i] = r > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)r;
}
} // namespace msc

uint64_t mscomp_gpu_xpress_huff_index_bound(uint64_t out_cap) { return xhi_max_segments(out_cap); }

static uint32_t xhd_grid(const mscomp_gpu_ctx* ctx, uint32_t count)
{
    const uint32_t groups = (count + 32 * XHD_WARPS_PER_BLOCK - 1) / (32 * XHD_WARPS_PER_BLOCK), cap = (uint32_t)ctx->sm_count * 2;
    return groups < cap ? groups : cap;
}

// Streams the indexed decoder could not follow (an index that does not describe the stream) are decoded by the lane-per-stream
// decoder, which needs no index and reproduces the reference's status codes; it finds nothing to do otherwise.
static void xhuff_second_chance(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b)
{
    unsigned long long* d_cnt = ctx->get_counters();
    MSC_OPT_IN_SMEM(xhuff_decode_marked_kernel, XHD_SMEM_BYTES);
    MSC_LAUNCH_SMEM(ctx, xhuff_decode_marked_kernel, xhd_grid(ctx, b->count), 32 * XHD_WARPS_PER_BLOCK, XHD_SMEM_BYTES,
                    d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status, (int)ST_INDEX,
                    d_cnt);
}

// Streams the indexed decoder handed to the lane-per-stream decoder since the context was created (diagnostic; synchronises).
uint64_t mscomp_gpu_xpress_huff_second_chances(mscomp_gpu_ctx* ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    unsigned long long* d_cnt = ctx->get_counters();
    if (!d_cnt) return 0;
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return h;
}

static uint32_t xhx_grid(const mscomp_gpu_ctx* ctx, uint32_t count)
{
    static int per_sm = 0;
    if (!per_sm) {
        cudaFuncSetAttribute(xhuff_decode_indexed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, xhuff_decode_indexed_kernel, XHX_WARPS_PER_BLOCK * 32, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 8; }
    }
    const uint32_t blocks = (count + XHX_WARPS_PER_BLOCK - 1) / XHX_WARPS_PER_BLOCK, cap = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
    return blocks < cap ? blocks : cap;
}

MSCompStatus mscomp_gpu_xpress_huff_index_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, const mscomp_gpu_batch* b,
                                                mscomp_gpu_xhuff_seg* d_index, const uint64_t* d_index_off, uint32_t* d_nseg)
{
    if (!ctx || !b || (b->count && (!d_in || !b->d_in_off || !b->d_in_len || !b->d_out_cap || !b->d_out_len || !b->d_status || !d_index || !d_index_off || !d_nseg)))
        return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    MSC_OPT_IN_SMEM(xhuff_index_kernel, XHD_SMEM_BYTES);
    MSC_LAUNCH_SMEM(ctx, xhuff_index_kernel, xhd_grid(ctx, b->count), 32 * XHD_WARPS_PER_BLOCK, XHD_SMEM_BYTES,
                    d_in, b->count, b->d_in_off, b->d_in_len, b->d_out_cap, (XhSeg*)d_index, d_index_off, d_nseg, b->d_out_len, b->d_status);
    return msc_cuda_status(cudaGetLastError(), "xpress_huff_index");
}

MSCompStatus mscomp_gpu_xpress_huff_decompress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b,
                                                             const mscomp_gpu_xhuff_seg* d_index, const uint64_t* d_index_off, const uint32_t* d_nseg)
{
    if (!batch_ok(ctx, d_in, d_out, b) || (b->count && (!d_index || !d_index_off || !d_nseg))) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    const uint32_t grid = xhx_grid(ctx, b->count);
    uint32_t* d_tok = (uint32_t*)ctx->get(SC_SLOTS, (size_t)grid * XHX_WARPS_PER_BLOCK * XHX_TOKBUF_WORDS * sizeof(uint32_t) + 64);
    if (!d_tok) return MSCOMP_MEM_ERROR;
    MSC_LAUNCH(ctx, xhuff_decode_indexed_kernel, grid, 32 * XHX_WARPS_PER_BLOCK,
               d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, (const XhSeg*)d_index, d_index_off, d_nseg,
               b->d_out_len, b->d_status, 0, d_tok);
    xhuff_second_chance(ctx, d_in, d_out, b);
    return msc_cuda_status(cudaGetLastError(), "xpress_huff_decompress");
}

MSCompStatus mscomp_gpu_xpress_huff_decompress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b)
{
    if (!batch_ok(ctx, d_in, d_out, b)) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    // Three ways to decode a batch that comes without an index (MSCOMP_LZ_DECODE selects one; default "tokens"):
    //   tokens  one lane per stream parses into 4-byte tokens (scratch: 4 bytes per output byte), one warp per stream executes them
    //   index   one lane per stream walks the stream and leaves a restart index, then one lane per SEGMENT decodes (the path of
    //           mscomp_gpu_xpress_huff_decompress_batch_indexed, which skips the walk when the caller kept the index)
    //   fused   one lane per stream does everything (no scratch at all; also the fallback when scratch cannot be had)
    // Measured on configs[3] (65,536 x 64 KiB): tokens 30.7 ms, index 15.7 + 18.5 ms, fused 40.7 ms -- the walk is the same serial
    // work in all three, so the index only pays when it is kept (18.5 ms).
    static const char* mode = getenv("MSCOMP_LZ_DECODE");
    const bool fused = mode && strcmp(mode, "fused") == 0;
    const bool use_index = mode && strcmp(mode, "index") == 0;
    if (!fused && !use_index) {
        MSC_OPT_IN_SMEM(xhuff_parse_kernel, XHD_SMEM_BYTES);
        uint32_t* tok = b->count >= 64 ? token_scratch(ctx, b) : nullptr;
        uint32_t* ntok = tok ? (uint32_t*)ctx->get(SC_AUX1, ((size_t)b->count + 1) * sizeof(uint32_t)) : nullptr;
        if (tok && ntok) {
            MSC_LAUNCH_SMEM(ctx, xhuff_parse_kernel, xhd_grid(ctx, b->count), 32 * XHD_WARPS_PER_BLOCK, XHD_SMEM_BYTES,
                            d_in, tok, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status, ntok);
            const uint32_t ab = (b->count + LZA_WARPS_PER_BLOCK - 1) / LZA_WARPS_PER_BLOCK, acap = (uint32_t)ctx->sm_count * 16;
            MSC_LAUNCH(ctx, lz_apply_kernel, ab < acap ? ab : acap, 32 * LZA_WARPS_PER_BLOCK, tok, ntok, b->d_out_off, b->d_status, b->count, d_out);
            return msc_cuda_status(cudaGetLastError(), "xpress_huff_decompress");
        }
    }
    // 1. room for the restart index of every stream -> offsets (one small kernel + the size scan), total to the host
    uint32_t* d_room = !use_index ? nullptr : (uint32_t*)ctx->get(SC_AUX1, ((size_t)b->count + 1) * sizeof(uint32_t) * 2);
    uint64_t* d_ioff = !use_index ? nullptr : (uint64_t*)ctx->get(SC_AUX0, ((size_t)b->count + 2) * sizeof(uint64_t));
    uint64_t* h_tot = (uint64_t*)ctx->get_pinned(256);
    XhSeg* d_index = nullptr;
    if (d_room && d_ioff && h_tot) {
        uint32_t* d_nseg = d_room + b->count + 1;
        MSC_LAUNCH(ctx, xhi_room_kernel, (b->count + 255) / 256, 256, b->d_in_len, b->d_out_cap, b->count, d_room);
        MSC_LAUNCH(ctx, scan_sizes_kernel, 1, SCAN_THREADS, (const uint32_t*)d_room, b->count, (uint64_t)0, (const uint64_t*)nullptr, d_ioff, h_tot + 8);
        if (cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
            const uint64_t total = h_tot[8];
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const uint64_t need = total * sizeof(XhSeg) + 64;
            if (need <= ctx->scratch_cap[SC_AUX2] || need <= (uint64_t)free_b / 2 + ctx->scratch_cap[SC_AUX2]) d_index = (XhSeg*)ctx->get(SC_AUX2, need);
        }
        const uint32_t grid = xhx_grid(ctx, b->count);
        uint32_t* d_tok = d_index ? (uint32_t*)ctx->get(SC_SLOTS, (size_t)grid * XHX_WARPS_PER_BLOCK * XHX_TOKBUF_WORDS * sizeof(uint32_t) + 64) : nullptr;
        if (d_index && d_tok) {
            // 2. the serial walk, one lane per stream: statuses, lengths and the index; 3. the decode proper, one lane per segment
            MSCompStatus st = mscomp_gpu_xpress_huff_index_batch(ctx, d_in, b, (mscomp_gpu_xhuff_seg*)d_index, d_ioff, d_nseg);
            if (st != MSCOMP_OK) return st;
            MSC_LAUNCH(ctx, xhuff_decode_indexed_kernel, grid, 32 * XHX_WARPS_PER_BLOCK,
                       d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, (const XhSeg*)d_index, (const uint64_t*)d_ioff,
                       (const uint32_t*)d_nseg, b->d_out_len, b->d_status, 1, d_tok);
            xhuff_second_chance(ctx, d_in, d_out, b);
            return msc_cuda_status(cudaGetLastError(), "xpress_huff_decompress");
        }
    }
    // no room for an index (or MSCOMP_LZ_DECODE=fused): every lane decodes its own stream and writes its own bytes
    MSC_OPT_IN_SMEM(xhuff_decode_kernel, XHD_SMEM_BYTES);
    MSC_LAUNCH_SMEM(ctx, xhuff_decode_kernel, xhd_grid(ctx, b->count), 32 * XHD_WARPS_PER_BLOCK, XHD_SMEM_BYTES,
                    d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status);
    return msc_cuda_status(cudaGetLastError(), "xpress_huff_decompress");
}

// Token execution of one chain of chunks by pointer jumping (kernels/lz_jump.cuh): d_ch = slot[nch] (u64) | ntok[nch] | pos[nch] as
// xhs_apply_chain_kernel takes them.  false: no room for the words (4 bytes per output byte) -- the caller executes in order.
static bool produce_by_jumping(mscomp_gpu_ctx* ctx, const uint32_t* d_tok, const uint8_t* d_ch, size_t nch, uint32_t produced, uint8_t* d_out)
{
    static const bool off = getenv("MSCOMP_LZ_JUMP") && strcmp(getenv("MSCOMP_LZ_JUMP"), "0") == 0;
    if (off || !produced) return false;
    const size_t bytes = (size_t)produced * sizeof(uint32_t) + 256;
    if (bytes > ctx->scratch_cap[SC_AUX0]) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (bytes > free_b / 2 + ctx->scratch_cap[SC_AUX0]) return false;
    }
    uint32_t* word = (uint32_t*)ctx->get(SC_AUX0, bytes);
    unsigned long long* ctr = ctx->get_counters();
    if (!word || !ctr) return false;
    uint32_t* unknown = (uint32_t*)(ctr + 8);
    const uint32_t cap = (uint32_t)ctx->sm_count * 8;
    const uint32_t blocks = (produced + LZJ_THREADS - 1) / LZJ_THREADS;
    MSC_LAUNCH(ctx, lzj_expand_kernel, (uint32_t)nch < cap ? (uint32_t)nch : cap, LZJ_THREADS, d_tok, (const uint64_t*)d_ch, (const uint32_t*)(d_ch + nch * 8),
               (const uint32_t*)(d_ch + nch * 12), (uint32_t)nch, word);
    uint32_t rounds = 1;
    while ((1ull << rounds) < produced) ++rounds;                                 // the distance to a chain's literal at least halves per round
    for (uint32_t r = 0; r < rounds; ++r) MSC_LAUNCH(ctx, lzj_round_kernel, blocks < cap * 2 ? blocks : cap * 2, LZJ_THREADS, word, produced);
    cudaMemsetAsync(unknown, 0, sizeof(uint32_t), ctx->stream);
    MSC_LAUNCH(ctx, lzj_store_kernel, blocks < cap * 2 ? blocks : cap * 2, LZJ_THREADS, (const uint32_t*)word, d_out, produced, unknown);
    return true;
}

// ---- one multi-chunk Xpress-Huffman stream: chunk-parallel parse (kernels/xhuff_chunks.cuh) --------------------------------
// d_in[0..in_len) is the stream, d_out has room for `cap` bytes.  *handled = true: the stream was decoded, *total bytes produced,
// status OK.  *handled = false: this path does not apply (one chunk, an incomplete code, a damaged stream, too little room or
// memory ...) and nothing is promised about d_out -- the caller runs the serial decoder, which owns the status codes.
MSCompStatus msc_xhuff_stream_decode(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint32_t in_len, uint8_t* d_out, uint64_t cap, uint64_t* total, bool* handled)
{
    *handled = false;
    if (!ctx || in_len < (uint32_t)XH_MIN_DATA + 2) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    // 1. candidate chunk starts
    const uint32_t capc = in_len / (uint32_t)XH_MIN_DATA + 2;
    uint32_t* d_cand = (uint32_t*)ctx->get(SC_AUX0, ((size_t)capc + 2) * sizeof(uint32_t));
    uint32_t* h_small = (uint32_t*)ctx->get_pinned(256);
    if (!d_cand || !h_small) return MSCOMP_OK;
    cudaMemsetAsync(d_cand + capc, 0, sizeof(uint32_t), ctx->stream);
    const uint32_t cblocks = (in_len + XHS_POS_PER_BLOCK - 1) / XHS_POS_PER_BLOCK, cmax = (uint32_t)ctx->sm_count * 8;
    MSC_LAUNCH(ctx, xhs_candidates_kernel, cblocks < cmax ? cblocks : cmax, XHS_THREADS, d_in, in_len, d_cand, d_cand + capc, capc);
    cudaMemcpyAsync(h_small + 32, d_cand + capc, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return MSCOMP_OK; }
    const uint32_t ncand = h_small[32];
    if (ncand < 1 || ncand > capc) return MSCOMP_OK;
    std::vector<uint32_t> cand(ncand);
    cudaMemcpyAsync(cand.data(), d_cand, (size_t)ncand * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return MSCOMP_OK; }
    std::sort(cand.begin(), cand.end());
    if (cand[0] != 0) return MSCOMP_OK;
    // 2. one lane per candidate parses one chunk into tokens.  A lone lane is latency-bound, so the candidates are spread over as
    //    many warps as the GPU holds (stride - 1 empty streams between two of them) instead of being packed 32 to a warp.
    // MSCOMP_XHUFF_CHUNK_PARSE=lane: the batch parser in speculative mode, one lane per candidate; default: a warp per candidate
    // (kernels/xhuff_chunk_warp.cuh: shared lookup table, lane 0 walks the symbols -- 5x fewer dependent instructions per symbol)
    static const bool lane_parse = getenv("MSCOMP_XHUFF_CHUNK_PARSE") && strcmp(getenv("MSCOMP_XHUFF_CHUNK_PARSE"), "lane") == 0;
    static const bool serial_parse = getenv("MSCOMP_XHUFF_CHUNK_PARSE") && strcmp(getenv("MSCOMP_XHUFF_CHUNK_PARSE"), "serial") == 0;   // lane 0 of the warp walks every symbol
    const uint32_t warps_avail = (uint32_t)ctx->sm_count * 2 * XHD_WARPS_PER_BLOCK;
    uint32_t per_warp = (ncand + warps_avail - 1) / warps_avail;
    uint32_t stride = lane_parse ? 32 : 1;
    while (stride > 1 && 32 / stride < per_warp) stride >>= 1;
    const uint64_t count64 = (uint64_t)ncand * stride;
    const uint64_t tok_bytes = (uint64_t)ncand * XH_CHUNK * sizeof(uint32_t);
    if (count64 > 0x7FFFFFFFull) return MSCOMP_OK;
    {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (tok_bytes + 64 > ctx->scratch_cap[SC_AUX2] && tok_bytes > (uint64_t)free_b / 2 + ctx->scratch_cap[SC_AUX2]) return MSCOMP_OK;
    }
    const uint32_t count = (uint32_t)count64;
    // descriptor block: in_off[count] out_off[count] (u64) | in_len out_cap out_len status ntok end_ip reach [count] (u32)
    const size_t desc_bytes = (size_t)count * (8 + 8 + 7 * 4);
    uint8_t* d_desc = (uint8_t*)ctx->get(SC_AUX1, desc_bytes + 64);
    uint8_t* h_desc = (uint8_t*)ctx->get_pinned_big(desc_bytes + 64);
    uint32_t* d_tok = (uint32_t*)ctx->get(SC_AUX2, tok_bytes + 64);
    if (!d_desc || !h_desc || !d_tok) return MSCOMP_OK;
    uint64_t* h_in_off = (uint64_t*)h_desc; uint64_t* h_out_off = h_in_off + count;
    uint32_t* h_u32 = (uint32_t*)(h_out_off + count);
    uint32_t *h_in_len = h_u32, *h_out_cap = h_u32 + count, *h_out_len = h_u32 + 2 * (size_t)count, *h_ntok = h_u32 + 4 * (size_t)count,
             *h_end = h_u32 + 5 * (size_t)count, *h_reach = h_u32 + 6 * (size_t)count;
    int32_t* h_status = (int32_t*)(h_u32 + 3 * (size_t)count);
    memset(h_desc, 0, desc_bytes);
    for (uint32_t k = 0; k < ncand; ++k) {
        const size_t i = (size_t)k * stride;
        h_in_off[i] = cand[k]; h_in_len[i] = in_len - cand[k]; h_out_off[i] = (uint64_t)k * XH_CHUNK; h_out_cap[i] = XH_CHUNK;
    }
    cudaMemcpyAsync(d_desc, h_desc, desc_bytes, cudaMemcpyHostToDevice, ctx->stream);
    const uint64_t* d_in_off = (const uint64_t*)d_desc; const uint64_t* d_out_off = d_in_off + count;
    uint32_t* d_u32 = (uint32_t*)(d_out_off + count);
    if (lane_parse) {
        MSC_OPT_IN_SMEM(xhuff_parse_chunks_kernel, XHD_SMEM_BYTES);
        MSC_LAUNCH_SMEM(ctx, xhuff_parse_chunks_kernel, xhd_grid(ctx, count), 32 * XHD_WARPS_PER_BLOCK, XHD_SMEM_BYTES,
                        d_in, d_tok, count, d_in_off, (const uint32_t*)d_u32, d_out_off, (const uint32_t*)(d_u32 + count),
                        d_u32 + 2 * (size_t)count, (int32_t*)(d_u32 + 3 * (size_t)count), d_u32 + 4 * (size_t)count, d_u32 + 5 * (size_t)count,
                        d_u32 + 6 * (size_t)count);
    } else {
        const uint32_t wb = (count + XHW_WARPS_PER_BLOCK - 1) / XHW_WARPS_PER_BLOCK, wcap = (uint32_t)ctx->sm_count * 16;
        if (serial_parse)
            MSC_LAUNCH(ctx, xhuff_parse_chunks_warp_serial_kernel, wb < wcap ? wb : wcap, 32 * XHW_WARPS_PER_BLOCK,
                       d_in, d_tok, count, d_in_off, (const uint32_t*)d_u32, d_out_off, (const uint32_t*)(d_u32 + count),
                       d_u32 + 2 * (size_t)count, (int32_t*)(d_u32 + 3 * (size_t)count), d_u32 + 4 * (size_t)count, d_u32 + 5 * (size_t)count,
                       d_u32 + 6 * (size_t)count);
        else
            MSC_LAUNCH(ctx, xhuff_parse_chunks_warp_kernel, wb < wcap ? wb : wcap, 32 * XHW_WARPS_PER_BLOCK,
                       d_in, d_tok, count, d_in_off, (const uint32_t*)d_u32, d_out_off, (const uint32_t*)(d_u32 + count),
                       d_u32 + 2 * (size_t)count, (int32_t*)(d_u32 + 3 * (size_t)count), d_u32 + 4 * (size_t)count, d_u32 + 5 * (size_t)count,
                       d_u32 + 6 * (size_t)count);
    }
    cudaMemcpyAsync(h_out_len, d_u32 + 2 * (size_t)count, (size_t)count * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return MSCOMP_OK; }
    // 3. stitch the chain from offset 0
    std::vector<uint32_t> chain;                     // candidate numbers in stream order
    std::vector<uint32_t> pos;                       // output position of each chunk
    uint64_t produced = 0;
    uint32_t at = 0;                                 // stream offset of the chunk being linked
    bool ok = false, independent = true;
    for (;;) {
        const auto it = std::lower_bound(cand.begin(), cand.end(), at);
        if (it == cand.end() || *it != at) break;
        const uint32_t k = (uint32_t)(it - cand.begin());
        const size_t i = (size_t)k * stride;
        if (h_status[i] < 0 || (uint64_t)h_reach[i] > produced || produced > 0xFFFFFFFFull - XH_CHUNK) break;
        if (h_reach[i]) independent = false;
        chain.push_back(k); pos.push_back((uint32_t)produced);
        produced += h_out_len[i];
        if (h_status[i] == 0) { ok = true; break; }                              // the stream ended inside this chunk (:83)
        const uint64_t next = (uint64_t)at + h_end[i];
        if (next == in_len) { ok = true; break; }                                // a finished chunk at the end of the input (:145-149)
        if (next > in_len || in_len - next < (uint64_t)XH_MIN_DATA || h_out_len[i] == 0) break;
        at = (uint32_t)next;
    }
    {
        static const bool trace = getenv("MSCOMP_TRACE") != nullptr;
        if (trace) {
            uint32_t nreach = 0, maxreach = 0, first = 0xFFFFFFFFu;
            for (size_t j = 0; j < chain.size(); ++j) { const uint32_t r = h_reach[(size_t)chain[j] * stride]; if (r) { ++nreach; if (first == 0xFFFFFFFFu) first = (uint32_t)j; } if (r > maxreach) maxreach = r; }
            fprintf(stderr, "[mscomp] xhuff stream: %u candidates, %zu chunks linked, ok=%d, produced=%llu, chunks reaching back: %u (first %u, max %u bytes)\n",
                    ncand, chain.size(), (int)ok, (unsigned long long)produced, nreach, first, maxreach);
        }
    }
    if (!ok || produced > cap) return MSCOMP_OK;
    // 4. execute the tokens
    const size_t nch = chain.size();
    {
        const size_t cb = nch * (8 + 4 + 4);
        uint8_t* d_ch = (uint8_t*)ctx->get(SC_OFFS, cb + 64);
        if (!d_ch) return MSCOMP_OK;
        std::vector<uint32_t> nt(nch);
        for (size_t j = 0; j < nch; ++j) nt[j] = h_ntok[(size_t)chain[j] * stride];
        uint64_t* h_slot = (uint64_t*)h_desc;                                     // (the descriptors are not needed any more)
        uint32_t* h_nt = (uint32_t*)(h_slot + nch); uint32_t* h_pos = h_nt + nch;
        for (size_t j = 0; j < nch; ++j) { h_slot[j] = (uint64_t)chain[j] * XH_CHUNK; h_nt[j] = nt[j]; h_pos[j] = pos[j]; }
        cudaMemcpyAsync(d_ch, h_desc, cb, cudaMemcpyHostToDevice, ctx->stream);
        // Few chunks: pointer jumping even when no chunk reaches back -- a warp executing its chunk's ≈ 12,000 tokens in order takes
        // as long as the parse, the order-free rounds use the whole GPU.  Many chunks fill the GPU by themselves with O(n) work.
        const bool own_warps = independent && nch >= (size_t)ctx->sm_count * 8;
        if (independent && !own_warps && produced <= 0x7FFFFFFFull && produce_by_jumping(ctx, (const uint32_t*)d_tok, d_ch, nch, (uint32_t)produced, d_out)) {
        } else if (independent) {                                                 // every chunk on its own warp
            const uint32_t ab = ((uint32_t)nch + LZA_WARPS_PER_BLOCK - 1) / LZA_WARPS_PER_BLOCK, acap = (uint32_t)ctx->sm_count * 16;
            MSC_LAUNCH(ctx, xhs_apply_chunks_kernel, ab < acap ? ab : acap, 32 * LZA_WARPS_PER_BLOCK, (const uint32_t*)d_tok, (const uint64_t*)d_ch,
                       (const uint32_t*)(d_ch + nch * 8), (const uint32_t*)(d_ch + nch * 12), (uint32_t)nch, d_out);
        } else if (produced <= 0x7FFFFFFFull && produce_by_jumping(ctx, (const uint32_t*)d_tok, d_ch, nch, (uint32_t)produced, d_out)) {
            // matches reach across chunk starts: no order at all, pointer jumping over one word per output byte (lz_jump.cuh)
        } else {                                                                  // ... or, without room for the words, in stream order by one block
            MSC_OPT_IN_SMEM(xhs_apply_chain_kernel, LZR_RING);
            MSC_LAUNCH_SMEM(ctx, xhs_apply_chain_kernel, 1, 32 * LZB_WARPS, LZR_RING, (const uint32_t*)d_tok, (const uint64_t*)d_ch, (const uint32_t*)(d_ch + nch * 8),
                            (const uint32_t*)(d_ch + nch * 12), (uint32_t)nch, d_out);
        }
    }
    const MSCompStatus st = msc_cuda_status(cudaGetLastError(), "xpress_huff_decompress");
    if (st != MSCOMP_OK) return st;
    *total = produced;
    *handled = true;
    return MSCOMP_OK;
}

// ---- one plain Xpress stream: flag groups found by speculation (kernels/xpress_spec.cuh) ------------------------------------
// d_in[0..in_len) is the stream, d_out has room for `cap` bytes (d_out == nullptr: the context's SC_OUT scratch, grown to what the stream
// produces).  *handled = true: the stream was decoded, *total bytes produced,
// status OK.  *handled = false: this path does not apply (a damaged stream, too little room or memory ...) and nothing is promised
// about d_out -- the caller runs the serial decoder, which owns the status codes.
MSCompStatus msc_xpress_stream_decode(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint32_t in_len, uint8_t* d_out, uint64_t cap, uint64_t* total, bool* handled)
{
    *handled = false;
    if (!ctx || in_len < 64 || in_len > (1u << 28)) return MSCOMP_OK;              // (24 bytes of records per stream byte stay with the context)
    cudaSetDevice(ctx->device);
    const uint32_t npos = in_len - 3;
    const uint32_t nblk = (npos + XPS_BLK - 1) / XPS_BLK;
    // layout of the small tables: table[nblk * XPS_NENT] entry[nblk] half[nblk] (u32) | count[nblk] base[nblk] (u64) | summary | result[2] | unit descriptors
    const size_t t_words = (size_t)nblk * XPS_NENT + 2 * (size_t)nblk;
    const size_t t_bytes = ((t_words * 4 + 15) & ~(size_t)15) + 2 * (size_t)nblk * 8 + 256;
    unsigned long long* rec = (unsigned long long*)ctx->get(SC_AUX3, (size_t)npos * 3 * sizeof(unsigned long long) + 64);
    uint8_t* tb = (uint8_t*)ctx->get(SC_AUX4, t_bytes);
    uint8_t* h_box = (uint8_t*)ctx->get_pinned(1024);
    if (!rec || !tb || !h_box) return MSCOMP_OK;
    uint32_t* table = (uint32_t*)tb;
    uint32_t* entry = table + (size_t)nblk * XPS_NENT;
    uint32_t* entry_half = entry + nblk;
    unsigned long long* count = (unsigned long long*)(tb + ((t_words * 4 + 15) & ~(size_t)15));
    unsigned long long* base = count + nblk;
    XpsSummary* d_sum = (XpsSummary*)(base + nblk);
    uint32_t* d_result = (uint32_t*)(d_sum + 1);
    uint8_t* d_unit = (uint8_t*)(d_result + 4);                                   // in_off out_off (u64) | in_len out_cap (u32)
    XpsSummary* h_sum = (XpsSummary*)(h_box + 512);                               // (the first 256 bytes are the result mailbox of the wrappers)
    uint32_t* h_result = (uint32_t*)(h_box + 512 + 64);
    uint8_t* h_unit = h_box + 512 + 128;
    const uint32_t wide = (uint32_t)ctx->sm_count * 8;
    // 1. every group that could be, the blocks' transfer tables, the true entry states, the true groups counted
    {
        const uint64_t nstate = (uint64_t)npos * 3, gb = (nstate + 255) / 256;
        MSC_LAUNCH(ctx, xps_groups_kernel, gb < wide * 4 ? (uint32_t)gb : wide * 4, 256, d_in, in_len, rec);
        MSC_OPT_IN_SMEM(xps_blocks_kernel, XPS_NST * sizeof(uint32_t));
        const uint32_t bcap = (uint32_t)ctx->sm_count * 2;
        MSC_LAUNCH_SMEM(ctx, xps_blocks_kernel, nblk < bcap ? nblk : bcap, XPS_THREADS, XPS_NST * sizeof(uint32_t), (const unsigned long long*)rec, in_len, nblk, table);
        MSC_LAUNCH(ctx, xps_entries_kernel, 1, 32, (const uint32_t*)table, nblk, entry, entry_half, d_sum);
        MSC_LAUNCH(ctx, xps_walk_kernel<false>, nblk < wide * 4 ? nblk : wide * 4, 32, d_in, in_len, (const unsigned long long*)rec, (const uint32_t*)entry,
                   (const uint32_t*)entry_half, (const XpsSummary*)d_sum, count, (const unsigned long long*)base, (XpressSeg*)nullptr);
        MSC_LAUNCH(ctx, xps_scan_kernel, 1, 32, (const unsigned long long*)count, base, d_sum, (unsigned long long)cap);
        cudaMemcpyAsync(h_sum, d_sum, sizeof(XpsSummary), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return MSCOMP_OK; }
    }
    {
        static const bool trace = getenv("MSCOMP_TRACE") != nullptr;
        if (trace) fprintf(stderr, "[mscomp] xpress stream: %u blocks, ok=%u, %u groups, %u restart entries, %llu bytes\n", nblk, h_sum->ok, h_sum->groups,
                           h_sum->nseg, (unsigned long long)h_sum->bytes);
    }
    if (!h_sum->ok || !h_sum->nseg || h_sum->bytes > cap || h_sum->bytes > 0x7FFFFFFFull) return MSCOMP_OK;
    const uint32_t nseg = h_sum->nseg, produced = (uint32_t)h_sum->bytes;
    if (!produced) return MSCOMP_OK;
    // 2. the restart index, the indexed parse (one lane per segment), the check that the segments join up
    const size_t ix_bytes = ((size_t)nseg * sizeof(XpressSeg) + 15) & ~(size_t)15;
    uint8_t* ixb = (uint8_t*)ctx->get(SC_AUX1, ix_bytes + (size_t)nseg * 3 * sizeof(uint32_t) + 64);
    uint32_t* tok = (uint32_t*)ctx->get(SC_AUX2, ((size_t)produced + 64) * sizeof(uint32_t));
    uint8_t* d_ch = (uint8_t*)ctx->get(SC_OFFS, (size_t)nseg * 16 + 64);
    if (!ixb || !tok || !d_ch) return MSCOMP_OK;
    XpressSeg* index = (XpressSeg*)ixb;
    uint32_t* seg_ntok = (uint32_t*)(ixb + ix_bytes);
    int32_t* seg_status = (int32_t*)(seg_ntok + nseg);
    uint32_t* seg_end = seg_ntok + 2 * (size_t)nseg;
    {
        uint64_t* hu = (uint64_t*)h_unit; hu[0] = 0; hu[1] = 0;
        uint32_t* hw = (uint32_t*)(hu + 2); hw[0] = in_len; hw[1] = produced;     // capacity = what the stream produces: a segment that disagrees fails
        cudaMemcpyAsync(d_unit, h_unit, 24, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(d_result, 0, 16, ctx->stream);
        MSC_LAUNCH(ctx, xps_walk_kernel<true>, nblk < wide * 4 ? nblk : wide * 4, 32, d_in, in_len, (const unsigned long long*)rec, (const uint32_t*)entry,
                   (const uint32_t*)entry_half, (const XpsSummary*)d_sum, count, (const unsigned long long*)base, index);
        const uint32_t per_block = 32 * XPD_WARPS_PER_BLOCK;
        // a lone lane is latency-bound: one segment per warp would be ideal, but the kernel maps 32 consecutive segments to a warp
        const uint32_t groups = (nseg + per_block - 1) / per_block;
        MSC_LAUNCH(ctx, xpress_parse_indexed_kernel, groups < wide * 2 ? groups : wide * 2, per_block, d_in, tok, 1u, (const uint64_t*)d_unit,
                   (const uint32_t*)(d_unit + 16), (const uint64_t*)(d_unit + 8), (const uint32_t*)(d_unit + 20), (const XpressSeg*)index, nseg,
                   seg_ntok, seg_status, seg_end);
        const uint32_t cb = (nseg + 255) / 256;
        MSC_LAUNCH(ctx, xps_chain_kernel, cb < wide ? cb : wide, 256, (const XpressSeg*)index, nseg, (const uint32_t*)seg_ntok, (const int32_t*)seg_status,
                   (const uint32_t*)seg_end, (unsigned long long*)d_ch, (uint32_t*)(d_ch + (size_t)nseg * 8), (uint32_t*)(d_ch + (size_t)nseg * 12), d_result);
        cudaMemcpyAsync(h_result, d_result, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return MSCOMP_OK; }
    }
    if (h_result[0] != 0 || h_result[1] != produced) return MSCOMP_OK;
    // 3. the tokens, in no order at all (kernels/lz_jump.cuh)
    if (!d_out) d_out = (uint8_t*)ctx->get(SC_OUT, (size_t)produced + 64);        // (the one-shot wrapper: room for what the stream really produces)
    if (!d_out) return MSCOMP_OK;
    if (!produce_by_jumping(ctx, (const uint32_t*)tok, d_ch, nseg, produced, d_out)) return MSCOMP_OK;
    const MSCompStatus st = msc_cuda_status(cudaGetLastError(), "xpress_decompress");
    if (st != MSCOMP_OK) return st;
    *total = produced;
    *handled = true;
    return MSCOMP_OK;
}

// Plain Xpress encode of a batch; seg_tiles / index / stride describe the optional restart index.  Large buffers take the two-pass
// variant (match finding per 64 KiB segment on its own block, then emission per buffer); small ones the single-pass kernel.
static MSCompStatus xpress_encode_launch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b,
                                         uint32_t seg_tiles, XpressSeg* index, uint32_t index_stride)
{
    cudaSetDevice(ctx->device);
    MSC_OPT_IN_SMEM(xpress_encode_kernel, sizeof(XpEncSmem));
    const uint32_t cap = (uint32_t)ctx->sm_count * 4;
    // Default: the reference's parse (lz_exact.cuh), byte-identical to its xpress_compress.  MSCOMP_GPU_ENCODER_FAST selects the earlier
    // bulk-synchronous matcher (1 table candidate + 1 lane peer; valid streams, other bytes): fused ("one") for small buffers, per
    // 64 KiB segment ("two") for large ones; MSCOMP_XPRESS_ENCODE=one/two forces a variant of it.
    static const char* mode = getenv("MSCOMP_XPRESS_ENCODE");
    const bool fast = ctx->encoder == MSCOMP_GPU_ENCODER_FAST;
    const bool fast_one = fast && mode && strcmp(mode, "one") == 0;
    bool fast_two = fast && mode && strcmp(mode, "two") == 0;
    bool two_pass = false, exact = false;
    uint32_t* seg_first = nullptr;
    uint32_t* work = nullptr;
    unsigned long long total_segs = 0;
    if (!fast_one) {
        seg_first = (uint32_t*)ctx->get(SC_AUX1, ((size_t)b->count + 2) * sizeof(uint32_t) + 64);
        unsigned long long* d_tot = (unsigned long long*)ctx->get(SC_RESULT, 1024);
        unsigned long long* h_tot = (unsigned long long*)ctx->get_pinned(1024);
        if (seg_first && d_tot && h_tot) {
            d_tot += 64; h_tot += 64;                                  // (clear of the result mailbox at the front)
            work = (uint32_t*)(d_tot + 4);
            MSC_LAUNCH(ctx, xpress_segments_kernel, 1, 1024, b->d_in_len, b->count, seg_first, d_tot, work);
            cudaMemcpyAsync(h_tot, d_tot, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
                total_segs = h_tot[0];
                if (fast && !fast_two && total_segs && h_tot[1] / b->count >= (128u << 10)) fast_two = true;
                if (total_segs < (1ull << 24)) { two_pass = fast_two && total_segs > 0; exact = !fast; }
            }
        }
    }
    XpPre pre{};
    if (!fast && !exact) return MSCOMP_MEM_ERROR;                      // the reference parse never falls back to the other matcher
    if (exact) {
        const size_t npos = (size_t)total_segs * LZX_SEG, tiles = npos / LZF_TILE;
        const size_t lo_b = npos * 4, v_b = npos / 8, x_b = tiles * 4 + (npos / LZX_SUB) * 4;   // answers of Find, token starts, exits per tile + per sub-tile
        uint8_t* sc = (uint8_t*)ctx->get(SC_AUX2, lo_b + v_b + x_b + npos * 2 + 256);
        if (!sc) return MSCOMP_MEM_ERROR;
        pre.seg_first = seg_first;
        pre.lenoff = (uint32_t*)sc;
        pre.visited = (uint32_t*)(sc + lo_b);
        pre.next_p = (uint32_t*)(sc + lo_b + v_b);
        uint16_t* dist = (uint16_t*)(sc + lo_b + v_b + x_b);
        MSC_OPT_IN_SMEM(xpress_emit_exact_kernel, sizeof(XpEncSmem));
        if (total_segs) {
            const size_t prev_smem = 32768 * sizeof(uint16_t) + LZX_STAGE + 128;
            MSC_OPT_IN_SMEM(lz_prev_kernel<uint16_t>, prev_smem);
            const unsigned long long pieces = total_segs * (LZX_SEG / LZX_PIECE);
            const uint32_t pg = pieces < (unsigned long long)ctx->sm_count * 3 ? (uint32_t)pieces : (uint32_t)ctx->sm_count * 3;
            MSC_LAUNCH_SMEM(ctx, lz_prev_kernel<uint16_t>, pg, 32, prev_smem, d_in, b->count, b->d_in_off, b->d_in_len, (const uint32_t*)seg_first, work,
                            dist, 8192u, LZX_PIECE);
            const uint32_t fg = total_segs < (unsigned long long)ctx->sm_count * LZX_FIND_MINB ? (uint32_t)total_segs : (uint32_t)ctx->sm_count * LZX_FIND_MINB;
            cudaMemsetAsync(pre.lenoff, 0xFF, lo_b, ctx->stream);       // LZX_UNKNOWN: nobody has asked
            MSC_LAUNCH(ctx, lz_find_kernel<LZX_FIND_MINB>, fg, LZX_FIND_THREADS, d_in, b->count, b->d_in_off, b->d_in_len, (const uint32_t*)seg_first,
                       work + 1, (const uint16_t*)dist, pre.lenoff, pre.visited, pre.next_p + tiles, 8192u, 0u);
            const unsigned long long sb = (tiles + 127) / 128;
            const uint32_t qg = sb < (unsigned long long)ctx->sm_count * 16 ? (uint32_t)sb : (uint32_t)ctx->sm_count * 16;
            MSC_LAUNCH(ctx, lz_stitch_kernel, qg, 128, d_in, b->count, b->d_in_off, b->d_in_len, (const uint32_t*)seg_first, (const uint16_t*)dist,
                       pre.lenoff, pre.visited, (const uint32_t*)(pre.next_p + tiles), pre.next_p, 8192u, 0u);
        }
        pre.dist = dist;
        MSC_LAUNCH_SMEM(ctx, xpress_emit_exact_kernel, b->count < cap ? b->count : cap, XPE_THREADS, sizeof(XpEncSmem),
                        d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status,
                        seg_tiles, index, index_stride, pre);
        return msc_cuda_status(cudaGetLastError(), "xpress_compress");
    }
    if (two_pass) {
        const size_t tiles = (size_t)total_segs * XPE_SEG_TILES;
        const size_t lo_b = tiles * LZF_TILE * sizeof(uint32_t), v_b = tiles * (LZF_TILE / 32) * sizeof(uint32_t), n_b = tiles * sizeof(uint32_t);
        uint8_t* sc = (uint8_t*)ctx->get(SC_AUX2, lo_b + v_b + n_b + 256);
        if (!sc) two_pass = false;
        else {
            pre.seg_first = seg_first;
            pre.lenoff = (uint32_t*)sc;
            pre.visited = (uint32_t*)(sc + lo_b);
            pre.next_p = (uint32_t*)(sc + lo_b + v_b);
        }
    }
    if (two_pass) {
        MSC_OPT_IN_SMEM(xpress_match_kernel, sizeof(LzFrontSmem));
        MSC_OPT_IN_SMEM(xpress_emit_kernel, sizeof(XpEncSmem));
        const uint32_t mg = total_segs < (unsigned long long)ctx->sm_count * 5 ? (uint32_t)total_segs : (uint32_t)ctx->sm_count * 5;
        MSC_LAUNCH_SMEM(ctx, xpress_match_kernel, mg, XPE_THREADS, sizeof(LzFrontSmem), d_in, b->count, b->d_in_off, b->d_in_len,
                        (const uint32_t*)seg_first, work, pre.lenoff, pre.visited, pre.next_p);
        MSC_LAUNCH_SMEM(ctx, xpress_emit_kernel, b->count < cap ? b->count : cap, XPE_THREADS, sizeof(XpEncSmem),
                        d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status,
                        seg_tiles, index, index_stride, pre);
    } else {
        MSC_LAUNCH_SMEM(ctx, xpress_encode_kernel, b->count < cap ? b->count : cap, XPE_THREADS, sizeof(XpEncSmem),
                        d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status,
                        seg_tiles, index, index_stride);
    }
    return msc_cuda_status(cudaGetLastError(), "xpress_compress");
}

MSCompStatus mscomp_gpu_xpress_compress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b)
{
    if (!batch_ok(ctx, d_in, d_out, b)) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    return xpress_encode_launch(ctx, d_in, d_out, b, 0u, nullptr, 0u);
}

static_assert(sizeof(XpressSeg) == sizeof(mscomp_gpu_xpress_seg), "xpress index entry layout");

MSCompStatus mscomp_gpu_xpress_compress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b,
                                                      uint32_t seg_bytes, mscomp_gpu_xpress_seg* d_index, uint32_t index_stride)
{
    if (!batch_ok(ctx, d_in, d_out, b) || !d_index || !index_stride || seg_bytes < LZF_TILE || seg_bytes % LZF_TILE) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    return xpress_encode_launch(ctx, d_in, d_out, b, seg_bytes / LZF_TILE, (XpressSeg*)d_index, index_stride);
}

MSCompStatus mscomp_gpu_xpress_decompress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b)
{
    if (!batch_ok(ctx, d_in, d_out, b)) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    const uint32_t groups = (b->count + 32 * XPD_WARPS_PER_BLOCK - 1) / (32 * XPD_WARPS_PER_BLOCK), cap = (uint32_t)ctx->sm_count * 8;
    // Few streams (down to the single buffer of the one-shot call): a WARP parses a stream, 32 items at a time, and a BLOCK
    // applies it; many streams: a lane parses a stream and a warp applies it -- the batch itself fills the GPU then.
    static const char* pmode = getenv("MSCOMP_XPRESS_PARSE");         // "lane" / "warp" force one of the parsers
    static const char* amode = getenv("MSCOMP_LZ_APPLY");             // "warp" / "block" force one of the apply kernels
    const bool warp_parse = pmode ? strcmp(pmode, "warp") == 0 : b->count < (uint32_t)ctx->sm_count * 64;
    const bool block_apply = amode ? strcmp(amode, "block") == 0 : b->count < (uint32_t)ctx->sm_count * 24;
    uint32_t* tok = (warp_parse || b->count >= 64) ? token_scratch(ctx, b) : nullptr;
    uint32_t* ntok = tok ? (uint32_t*)ctx->get(SC_AUX1, ((size_t)b->count + 1) * sizeof(uint32_t)) : nullptr;
    if (tok && ntok) {
        if (warp_parse) {
            const uint32_t wb = (b->count + XPD_WARPS_PER_BLOCK - 1) / XPD_WARPS_PER_BLOCK, wcap = (uint32_t)ctx->sm_count * 16;
            MSC_LAUNCH(ctx, xpress_parse_warp_kernel, wb < wcap ? wb : wcap, 32 * XPD_WARPS_PER_BLOCK,
                       d_in, tok, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status, ntok);
        } else {
            MSC_LAUNCH(ctx, xpress_parse_kernel, groups < cap ? groups : cap, 32 * XPD_WARPS_PER_BLOCK,
                       d_in, tok, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status, ntok);
        }
        if (block_apply) {
            const uint32_t acap = (uint32_t)ctx->sm_count * 8;
            MSC_LAUNCH(ctx, lz_apply_block_kernel, b->count < acap ? b->count : acap, 32 * LZB_WARPS, tok, ntok, b->d_out_off, b->d_status, b->count, d_out);
        } else {
            const uint32_t ab = (b->count + LZA_WARPS_PER_BLOCK - 1) / LZA_WARPS_PER_BLOCK, acap = (uint32_t)ctx->sm_count * 16;
            MSC_LAUNCH(ctx, lz_apply_kernel, ab < acap ? ab : acap, 32 * LZA_WARPS_PER_BLOCK, tok, ntok, b->d_out_off, b->d_status, b->count, d_out);
        }
    } else {
        MSC_LAUNCH(ctx, xpress_decode_kernel, groups < cap ? groups : cap, 32 * XPD_WARPS_PER_BLOCK,
                   d_in, d_out, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, b->d_out_len, b->d_status);
    }
    return msc_cuda_status(cudaGetLastError(), "xpress_decompress");
}

MSCompStatus mscomp_gpu_xpress_decompress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* b,
                                                        const mscomp_gpu_xpress_seg* d_index, uint32_t index_stride)
{
    if (!batch_ok(ctx, d_in, d_out, b) || !d_index || !index_stride) return MSCOMP_ARG_ERROR;
    if (!b->count) return MSCOMP_OK;
    cudaSetDevice(ctx->device);
    const uint64_t lanes = (uint64_t)b->count * index_stride;
    uint32_t* tok = lanes < 0x7fffffffu ? token_scratch(ctx, b) : nullptr;
    uint32_t* seg = tok ? (uint32_t*)ctx->get(SC_AUX1, (size_t)lanes * 3 * sizeof(uint32_t) + 64) : nullptr;
    if (!tok || !seg) return mscomp_gpu_xpress_decompress_batch(ctx, d_in, d_out, b);     // no room for tokens: one lane per stream
    uint32_t* seg_ntok = seg;
    int32_t* seg_status = (int32_t*)(seg + lanes);
    uint32_t* seg_end = seg + 2 * lanes;
    const uint32_t per_block = 32 * XPD_WARPS_PER_BLOCK;
    const uint64_t groups = (lanes + per_block - 1) / per_block;
    const uint32_t cap = (uint32_t)ctx->sm_count * 16;
    MSC_LAUNCH(ctx, xpress_parse_indexed_kernel, groups < cap ? (uint32_t)groups : cap, per_block,
               d_in, tok, b->count, b->d_in_off, b->d_in_len, b->d_out_off, b->d_out_cap, (const XpressSeg*)d_index, index_stride,
               seg_ntok, seg_status, seg_end);
    // few streams: a block per stream (rounds pipelined over its warps); many streams: a warp per stream fills the GPU by itself
    static const char* amode = getenv("MSCOMP_LZ_APPLY");             // "warp" / "block" force one of them
    const bool block_apply = amode ? strcmp(amode, "block") == 0 : b->count < (uint32_t)ctx->sm_count * 24;
    if (block_apply) {
        const uint32_t acap = (uint32_t)ctx->sm_count * 8;           // 8 blocks of 8 warps fill an SM
        MSC_LAUNCH(ctx, xpress_apply_indexed_block_kernel, b->count < acap ? b->count : acap, 32 * LZB_WARPS,
                   (const uint32_t*)tok, (const uint32_t*)seg_ntok, (const int32_t*)seg_status, (const uint32_t*)seg_end,
                   (const XpressSeg*)d_index, index_stride, b->d_out_off, b->count, d_out, b->d_out_len, b->d_status);
    } else {
        const uint32_t ab = (b->count + LZA_WARPS_PER_BLOCK - 1) / LZA_WARPS_PER_BLOCK, acap = (uint32_t)ctx->sm_count * 16;
        MSC_LAUNCH(ctx, xpress_apply_indexed_kernel, ab < acap ? ab : acap, 32 * LZA_WARPS_PER_BLOCK,
                   (const uint32_t*)tok, (const uint32_t*)seg_ntok, (const int32_t*)seg_status, (const uint32_t*)seg_end,
                   (const XpressSeg*)d_index, index_stride, b->d_out_off, b->count, d_out, b->d_out_len, b->d_status);
    }
    return msc_cuda_status(cudaGetLastError(), "xpress_decompress");
}

} // extern "C"
