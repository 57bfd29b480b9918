// mscomp_shim.cu -- the reference's host-pointer C ABI on top of the device batch API.
//
// Entry points and the reference code each one replaces:
//   lznt1_compress / lznt1_decompress / lznt1_max_compressed_size ....... src/lznt1_compress.cpp:27,:233-273,
//                                                                        src/lznt1_decompress.cpp:337 (+ :122-273 driver)
//   xpress_* ............................................................ src/xpress_compress.cpp:30,:240-347, src/xpress_decompress.cpp:405-462
//   xpress_huff_* ....................................................... src/xpress_huff_compress.cpp:46,:247-331, src/xpress_huff_decompress.cpp:136-162
//   ms_compress / ms_decompress / ms_max_compressed_size ................ src/mscomp.cpp:96-132 (+ the MSCOMP_NONE copy engine :26-32)
//
// Each call: host -> device copy of the input, kernels on the calling thread's stream, device -> host copy of the
// result.  Host-side work is limited to what the format forces to be serial on an unindexed stream (the LZNT1
// header walk) and to reproducing the reference's status codes.  No codec work happens on the CPU.
#include "ctx.cuh"
#include "shim_common.h"
#include "lznt1_host.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>

// ---- per-thread default context -------------------------------------------------------------------------------
mscomp_gpu_ctx* msc_thread_ctx()
{
    // Deliberately leaked at thread/process exit: CUDA may already be shut down when TLS destructors run.
    static thread_local mscomp_gpu_ctx* ctx = nullptr;
    static thread_local bool failed = false;
    if (ctx || failed) return ctx;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    ctx = mscomp_gpu_create(dev);
    if (!ctx) {
        failed = true;
        fprintf(stderr, "libMSCompression(B200): no usable CUDA device -- this library has no CPU implementation\n");
    }
    return ctx;
}

static inline unsigned rd16(const uint8_t* p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }

// Tuning / test hooks of the pipelined host wrappers (bytes): MSCOMP_PIPE_SLICE = upload/encode/download granularity of
// lznt1_compress, MSCOMP_PIPE_MIN = smallest stream lznt1_decompress pipelines, MSCOMP_PIPE_RANGE = its range size.
static size_t env_bytes(const char* name, size_t dflt)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const unsigned long long x = strtoull(v, nullptr, 10);
    return x ? (size_t)x : dflt;
}

namespace {
struct StreamGuard {            // the user's buffers must not be in use by a DMA once we return
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};
    ~StreamGuard() { for (cudaStream_t x : s) if (x) cudaStreamSynchronize(x); }
};
} // namespace

// =================================================================================================================
// LZNT1
// =================================================================================================================
extern "C" size_t lznt1_max_compressed_size(size_t in_len) { return in_len + 3 + 2 * ((in_len + 4095) / 4096); }

extern "C" MSCompStatus lznt1_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || !out) return MSCOMP_ARG_ERROR;
    const size_t cap = *out_len;
    if (in_len == 0) { *out_len = 0; return MSCOMP_OK; }   // no chunk, no terminator needed (src/lznt1_compress.cpp:239,:268-272)
    if (cap == 0) return MSCOMP_BUF_ERROR;                 // undefined upstream (src/lznt1_compress.cpp:239), rejected here
    mscomp_gpu_ctx* ctx = msc_thread_ctx();
    if (!ctx) return MSCOMP_MEM_ERROR;
    const size_t nchunks = (in_len + 4095) / 4096;
    const size_t dcap = in_len + 2 * nchunks + 16;
    uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + 64);
    uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, dcap + 64);
    mscomp_gpu_result* d_res = (mscomp_gpu_result*)ctx->get(SC_RESULT, 256);
    mscomp_gpu_result* h_res = (mscomp_gpu_result*)ctx->get_pinned(256);
    if (!d_in || !d_out || !d_res || !h_res) return MSCOMP_MEM_ERROR;
    MSCompStatus st;
    static const size_t kSliceEnv = env_bytes("MSCOMP_PIPE_SLICE", (size_t)64 << 20);
    const size_t kSlice = (kSliceEnv + 4095) / 4096 * 4096;  // upload / encode / download granularity (a multiple of the chunk size)
    if (in_len > kSlice && ctx->copy_ready()) {
        // Chunks are independent and the stream is their concatenation, so the call runs as a three-stage pipeline over
        // slices: upload slice k+1 (copy stream) | encode + compact slice k (compute stream) | download the bytes of slice
        // k-1 (download stream).  The device keeps the running end offset (scan_sizes_kernel continues from the previous
        // slice); the host learns it from a pinned mailbox one slice behind, to address the download.
        uint8_t* d_slots = (uint8_t*)ctx->get(SC_SLOTS, nchunks * MSC_LZNT1_SLOT_BYTES + 16);
        uint32_t* d_sizes = (uint32_t*)ctx->get(SC_SIZES, (nchunks + 1) * sizeof(uint32_t));
        uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (nchunks + 1) * sizeof(uint64_t));
        uint64_t* h_end = (uint64_t*)ctx->get_pinned(1024);
        if (!d_slots || !d_sizes || !d_offs || !h_end) return MSCOMP_MEM_ERROR;
        h_end += 32;                                                         // the first 256 bytes are the result mailbox
        StreamGuard guard;
        guard.s[0] = ctx->copy_stream; guard.s[1] = ctx->down_stream; guard.s[2] = ctx->stream;
        cudaEventRecord(ctx->copy_ev[15], ctx->stream);                      // earlier work on the scratch buffers
        cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[15], 0);
        cudaStreamWaitEvent(ctx->down_stream, ctx->copy_ev[15], 0);
        const size_t nslices = (in_len + kSlice - 1) / kSlice;
        size_t sent = 0;                                                     // compressed bytes already on their way to `out`
        bool overflow = false;
        auto download = [&](size_t k) -> MSCompStatus {                      // slice k is encoded and compacted: ship its bytes
            MSCompStatus s2 = msc_cuda_status(cudaEventSynchronize(ctx->work_ev[k % 15]), "lznt1_compress");
            if (s2 != MSCOMP_OK) return s2;
            const size_t end = (size_t)h_end[k % 15];
            if (end > cap) { overflow = true; return MSCOMP_OK; }
            if (end > sent && !overflow) cudaMemcpyAsync(out + sent, d_out + sent, end - sent, cudaMemcpyDeviceToHost, ctx->down_stream);
            sent = end;
            return MSCOMP_OK;
        };
        for (size_t k = 0; k < nslices; ++k) {
            const size_t off = k * kSlice, len = in_len - off < kSlice ? in_len - off : kSlice;
            const uint32_t first = (uint32_t)(off / 4096), cnt = (uint32_t)((len + 4095) / 4096);
            cudaEvent_t ev = ctx->copy_ev[k % 15];
            cudaMemcpyAsync(d_in + off, in + off, len, cudaMemcpyHostToDevice, ctx->copy_stream);
            cudaEventRecord(ev, ctx->copy_stream);
            cudaStreamWaitEvent(ctx->stream, ev, 0);
            if ((st = msc_lznt1_encode_chunks(ctx, d_in, in_len, first, cnt, d_slots, d_sizes)) != MSCOMP_OK) return st;
            // (the running end offset goes straight into the pinned mailbox: a small device-to-host memcpy here would
            // queue behind the bulk downloads on the copy engine and stall this stream)
            if ((st = msc_compact_slots_range(ctx, d_slots, MSC_LZNT1_SLOT_BYTES, d_sizes, first, cnt, d_offs, d_out, &h_end[k % 15])) != MSCOMP_OK) return st;
            cudaEventRecord(ctx->work_ev[k % 15], ctx->stream);
            if (k >= 1 && (st = download(k - 1)) != MSCOMP_OK) return st;   // one slice behind: the GPU always has slice k queued
        }
        if ((st = download(nslices - 1)) != MSCOMP_OK) return st;
        if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->down_stream), "lznt1_compress")) != MSCOMP_OK) return st;
        if (overflow) return MSCOMP_BUF_ERROR;                               // same outcome as the chunk-by-chunk checks upstream (:239-259)
        const size_t total = sent;
        if (cap - total >= 2) out[total] = out[total + 1] = 0;               // end-of-buffer terminator, not counted (:268-271)
        *out_len = total;
        return MSCOMP_OK;
    }
    if (in_len) cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
    if ((st = mscomp_gpu_lznt1_compress(ctx, d_in, in_len, d_out, dcap, nullptr, d_res)) != MSCOMP_OK) return st;
    cudaMemcpyAsync(h_res, d_res, sizeof(*h_res), cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_compress")) != MSCOMP_OK) return st;
    const size_t total = (size_t)h_res->bytes;
    if (total > cap) return MSCOMP_BUF_ERROR;              // same outcome as the chunk-by-chunk checks upstream (:239-259)
    if (total) cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_compress")) != MSCOMP_OK) return st;
    if (cap - total >= 2) out[total] = out[total + 1] = 0; // end-of-buffer terminator, not counted (:268-271)
    *out_len = total;
    return MSCOMP_OK;
}

using namespace msc_host;

static unsigned long long g_pipeline_fallbacks = 0;
// Diagnostic (not part of the reference ABI): how often the pipelined lznt1_decompress handed a stream to the simple path.
extern "C" MSCOMPAPI unsigned long long msc_pipeline_fallbacks(void) { return g_pipeline_fallbacks; }

namespace {
constexpr MSCompStatus MSC_FALLBACK = (MSCompStatus)1000;    // internal: "take the simple path"

// lznt1_decompress for large streams, as a pipeline over ranges of the compressed stream:
//   upload range k+1 (copy stream) | index + decode range k (compute stream) | download the output of range k-1.
// The chunk index is built on the GPU (kernels/lznt1_index.cuh) range by range, so nothing waits for a host-side walk
// over the headers (one cache miss per chunk).  This path handles the normal case only -- every chunk decodes, every
// chunk but the last is 4096 bytes, the output fits; anything else drains the pipeline and returns MSC_FALLBACK, and
// the caller runs the simple path, which reproduces the reference's status codes chunk by chunk.
MSCompStatus lznt1_decompress_pipelined(mscomp_gpu_ctx* ctx, const_bytes in, size_t in_len, bytes out, size_t* out_len, size_t range_bytes)
{
    const size_t cap = *out_len;
    const size_t S = MSC_LZNT1_INDEX_SEG;
    const size_t R = (range_bytes + S - 1) / S * S;
    const size_t nranges = (in_len + R - 1) / R;
    const size_t kOverlap = 8192;                                // a chunk that starts inside a range ends < 4098 bytes past it
    size_t ncap = cap / 4096 + 2;                                // more chunks than this cannot fit the output anyway
    if (ncap > in_len / 3 + 1) ncap = in_len / 3 + 1;
    if (ncap > 0x7fffffffu) return MSC_FALLBACK;
    uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + kOverlap + 64);
    uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, ncap * 4096 + 64);
    uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (ncap + 1) * sizeof(uint64_t));
    uint32_t* d_sizes = (uint32_t*)ctx->get(SC_SIZES, (ncap + 1) * sizeof(uint32_t));
    int32_t* d_status = (int32_t*)ctx->get(SC_STATUS, (ncap + 1) * sizeof(int32_t));
    uint8_t* d_box = (uint8_t*)ctx->get(SC_RESULT, 1024);        // index state at +0, per-range summaries at +256
    uint8_t* h_box = (uint8_t*)ctx->get_pinned(4096);            // per-range copies of the state at +256, of the summaries at +2048
    if (!d_in || !d_out || !d_offs || !d_sizes || !d_status || !d_box || !h_box) return MSC_FALLBACK;   // (an oversized *out_len makes this
                                                                         // path ask for far more than the stream needs; the simple path sizes exactly)
    msc_lznt1_index_state* d_state = (msc_lznt1_index_state*)d_box;
    auto h_state = [&](size_t k) { return (msc_lznt1_index_state*)(h_box + 256 + 64 * (k % 8)); };
    auto h_sum = [&](size_t k) { return (mscomp_gpu_result*)(h_box + 2048 + 64 * (k % 8)); };

    StreamGuard guard;
    guard.s[0] = ctx->copy_stream; guard.s[1] = ctx->down_stream; guard.s[2] = ctx->stream;
    cudaEventRecord(ctx->copy_ev[15], ctx->stream);              // earlier work on the scratch buffers
    cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[15], 0);
    cudaStreamWaitEvent(ctx->down_stream, ctx->copy_ev[15], 0);
    cudaMemsetAsync(d_state, 0, sizeof(*d_state), ctx->stream);

    static const bool trace = getenv("MSCOMP_TRACE") != nullptr;  // host-side timeline of the pipeline on stderr
    struct timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto tr = [&](const char* what, size_t k) {
        if (!trace) return;
        struct timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[lznt1_decompress] %8.1f us  %s %zu\n", (t.tv_sec - ts0.tv_sec) * 1e6 + (t.tv_nsec - ts0.tv_nsec) * 1e-3, what, k);
    };
    MSCompStatus st = MSCOMP_OK;
    auto stage_in = [&](size_t k) -> MSCompStatus {              // upload range k, then count its chunks
        tr("in", k);
        const size_t off = k * R;
        size_t len = in_len - off < R + kOverlap ? in_len - off : R + kOverlap;
        cudaMemcpyAsync(d_in + off, in + off, len, cudaMemcpyHostToDevice, ctx->copy_stream);
        cudaEventRecord(ctx->copy_ev[k % 15], ctx->copy_stream);
        cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k % 15], 0);
        const size_t bytes = in_len - off < R ? in_len - off : R;
        MSCompStatus s2 = msc_lznt1_index_count(ctx, d_in, in_len, (uint32_t)(off / S), (uint32_t)((bytes + S - 1) / S), (uint32_t)(R / S), (int)(k % 8), d_state, h_state(k));
        if (s2 != MSCOMP_OK) return s2;               // (state and summaries are written to the pinned mailbox by the kernels themselves:
                                                      // small device-to-host memcpys would queue behind the bulk downloads on the copy engine)
        cudaEventRecord(ctx->work_ev[(2 * k) % 16], ctx->stream);
        return MSCOMP_OK;
    };
    struct Range { size_t c0 = 0, n = 0; bool last = false; };
    std::vector<Range> rg(nranges);
    bool seen_end = false;
    auto stage_decode = [&](size_t k) -> MSCompStatus {          // chunk offsets, decode, summary of range k
        tr("wait-count", k);
        MSCompStatus s2 = msc_cuda_status(cudaEventSynchronize(ctx->work_ev[(2 * k) % 16]), "lznt1_decompress");
        if (s2 != MSCOMP_OK) return s2;
        tr("counted", k);
        const msc_lznt1_index_state hs = *h_state(k);
        if (hs.nchunks > ncap) return MSC_FALLBACK;
        rg[k].n = (size_t)hs.range_chunks;
        rg[k].c0 = (size_t)(hs.nchunks - hs.range_chunks);
        rg[k].last = hs.done && !seen_end;                       // the chain ends in this range: its last chunk may be short
        if (hs.done) seen_end = true;
        if (rg[k].n) {
            const size_t off = k * R, bytes = in_len - off < R ? in_len - off : R;
            if ((s2 = msc_lznt1_index_write(ctx, d_in, in_len, (uint32_t)(off / S), (uint32_t)((bytes + S - 1) / S), (uint32_t)(R / S), (int)(k % 8), d_offs + rg[k].c0)) != MSCOMP_OK) return s2;
            if ((s2 = msc_lznt1_decode_chunks(ctx, d_in, in_len, d_offs + rg[k].c0, (uint32_t)rg[k].n, d_out + rg[k].c0 * 4096,
                                              d_sizes + rg[k].c0, d_status + rg[k].c0)) != MSCOMP_OK) return s2;
            if ((s2 = msc_summarize_units(ctx, d_sizes + rg[k].c0, d_status + rg[k].c0, (uint32_t)rg[k].n, 4096, h_sum(k))) != MSCOMP_OK) return s2;
        }
        cudaEventRecord(ctx->work_ev[(2 * k + 1) % 16], ctx->stream);
        return MSCOMP_OK;
    };
    size_t total = 0, last_size = 4096;
    auto stage_out = [&](size_t k) -> MSCompStatus {             // range k is decoded: check it, ship its bytes
        tr("wait-decode", k);
        MSCompStatus s2 = msc_cuda_status(cudaEventSynchronize(ctx->work_ev[(2 * k + 1) % 16]), "lznt1_decompress");
        tr("decoded", k);
        if (s2 != MSCOMP_OK || !rg[k].n) return s2;
        const mscomp_gpu_result r = *h_sum(k);
        if (r.status != 0 || r.short_units != 0) return MSC_FALLBACK;
        const size_t last = (size_t)r.bytes - (rg[k].n - 1) * 4096;
        if (!rg[k].last && last != 4096) return MSC_FALLBACK;   // a short chunk in mid-stream
        if (rg[k].c0 * 4096 != total || (size_t)r.bytes > cap - total) return MSC_FALLBACK;
        cudaMemcpyAsync(out + total, d_out + total, (size_t)r.bytes, cudaMemcpyDeviceToHost, ctx->down_stream);
        total += (size_t)r.bytes;
        last_size = last;
        return MSCOMP_OK;
    };

    // uploads and chunk counts run `ahead` ranges in front of the decoder, so the copy engine never waits for the host
    static const size_t kAheadEnv = env_bytes("MSCOMP_PIPE_AHEAD", 1);
    const size_t ahead = kAheadEnv < 1 ? 1 : (kAheadEnv > 6 ? 6 : kAheadEnv);
    for (size_t k = 0; k < ahead && k < nranges && st == MSCOMP_OK; ++k) st = stage_in(k);
    for (size_t k = 0; k < nranges && st == MSCOMP_OK; ++k) {
        if (k + ahead < nranges) st = stage_in(k + ahead);
        if (st == MSCOMP_OK) st = stage_decode(k);
        if (st == MSCOMP_OK && k >= 1) st = stage_out(k - 1);
    }
    if (st == MSCOMP_OK) st = stage_out(nranges - 1);
    if (st != MSCOMP_OK) return st;                              // (the guard drains the streams)
    tr("wait-downloads", nranges);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->down_stream), "lznt1_decompress")) != MSCOMP_OK) return st;
    tr("done", nranges);
    const msc_lznt1_index_state fin = *h_state(nranges - 1);
    if (!fin.done || fin.nchunks == 0) return MSC_FALLBACK;
    const size_t nchunks = (size_t)fin.nchunks;
    size_t replayed = 0;
    st = lznt1_replay(nchunks, cap, (Lznt1Tail)fin.tail, [&](size_t c) { return c + 1 < nchunks ? (size_t)4096 : last_size; },
                      [](size_t) { return 0; }, &replayed);
    if (st == MSCOMP_OK) *out_len = replayed;
    return st;
}
} // namespace

extern "C" MSCompStatus lznt1_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || (!out && *out_len)) return MSCOMP_ARG_ERROR;
    const size_t cap = *out_len;

    mscomp_gpu_ctx* ctx = in_len >= 2 ? msc_thread_ctx() : nullptr;
    if (in_len >= 2 && !ctx) return MSCOMP_MEM_ERROR;

    // Large streams: pipelined, chunk index built on the GPU.  Whatever that path does not cover comes back as
    // MSC_FALLBACK and takes the simple path below.
    static const size_t kPipeMin = env_bytes("MSCOMP_PIPE_MIN", (size_t)4 << 20), kPipeRange = env_bytes("MSCOMP_PIPE_RANGE", 0);
    if (in_len >= kPipeMin && cap && ctx->copy_ready()) {
        size_t range = kPipeRange;
        if (!range) {                                            // about 8 ranges per call, 4..32 MiB each
            range = in_len / 8;
            if (range < ((size_t)4 << 20)) range = (size_t)4 << 20;
            if (range > ((size_t)32 << 20)) range = (size_t)32 << 20;
        }
        const MSCompStatus pst = lznt1_decompress_pipelined(ctx, in, in_len, out, out_len, range);
        if (pst != MSC_FALLBACK) return pst;
        __atomic_fetch_add(&g_pipeline_fallbacks, 1ull, __ATOMIC_RELAXED);
        *out_len = cap;
    }

    // Simple path.  The upload of the stream starts first, on the copy stream, so that it overlaps with the header walk below.
    uint8_t* d_in = ctx ? (uint8_t*)ctx->get(SC_IN, in_len + 64) : nullptr;
    if (ctx && !d_in) return MSCOMP_MEM_ERROR;
    struct CopyGuard {          // the user's buffer must not be in use by a DMA once we return
        cudaStream_t s = nullptr;
        ~CopyGuard() { if (s) cudaStreamSynchronize(s); }
    } guard;
    const bool early = ctx && in_len > ((size_t)1 << 20) && ctx->copy_ready();
    if (early) {
        cudaEventRecord(ctx->copy_ev[15], ctx->stream);                      // earlier work on the scratch buffer
        cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[15], 0);
        cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->copy_stream);
        cudaEventRecord(ctx->copy_ev[0], ctx->copy_stream);
        guard.s = ctx->copy_stream;
    }

    // Header walk: the format stores no chunk index and each header gives only the distance to the next one
    // (src/lznt1_decompress.cpp:127-135), so an unindexed stream is a linked list; it is walked here, on the host,
    // where the stream already is.
    std::vector<uint64_t> offs;
    offs.reserve(in_len / 2048 + 4);
    size_t ip = 0;
    const Lznt1Tail tail = lznt1_walk(in, in_len, offs, &ip);
    const size_t nchunks = offs.size();
    if (nchunks > 0x7fffffffu) return MSCOMP_ARG_ERROR;

    size_t total = 0;
    if (nchunks == 0) {
        MSCompStatus st = lznt1_replay(0, cap, tail, [](size_t) { return (size_t)0; }, [](size_t) { return 0; }, &total);
        if (st == MSCOMP_OK) *out_len = 0;
        return st;
    }

    uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, nchunks * 4096 + 64);
    uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (nchunks + 1) * sizeof(uint64_t));
    uint32_t* d_sizes = (uint32_t*)ctx->get(SC_SIZES, (nchunks + 1) * sizeof(uint32_t));
    int32_t* d_status = (int32_t*)ctx->get(SC_STATUS, (nchunks + 1) * sizeof(int32_t));
    mscomp_gpu_result* d_res = (mscomp_gpu_result*)ctx->get(SC_RESULT, 256);
    mscomp_gpu_result* h_res = (mscomp_gpu_result*)ctx->get_pinned(256);
    if (!d_out || !d_offs || !d_sizes || !d_status || !d_res || !h_res) return MSCOMP_MEM_ERROR;
    if (early) cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[0], 0);
    else cudaMemcpyAsync(d_in, in, ip, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_offs, offs.data(), nchunks * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    MSCompStatus st = mscomp_gpu_lznt1_decompress(ctx, d_in, ip, d_offs, (uint32_t)nchunks, d_out, d_sizes, d_status, d_res);
    if (st != MSCOMP_OK) return st;
    cudaMemcpyAsync(h_res, d_res, sizeof(*h_res), cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_decompress")) != MSCOMP_OK) return st;

    const mscomp_gpu_result r = *h_res;
    const uint8_t* d_final = d_out;
    if (r.status == 0 && r.short_units == 0) {
        // every chunk but the last produced 4096 bytes: sizes are implied, output is already contiguous
        const size_t last = (size_t)r.bytes - (nchunks - 1) * 4096;
        st = lznt1_replay(nchunks, cap, tail, [&](size_t c) { return c + 1 < nchunks ? (size_t)4096 : last; }, [](size_t) { return 0; }, &total);
    } else {
        // rare: a failing chunk, or short chunks in mid-stream (streams cut by a flushing compressor)
        std::vector<uint32_t> sizes(nchunks);
        std::vector<int32_t> status(nchunks);
        cudaMemcpyAsync(sizes.data(), d_sizes, nchunks * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(status.data(), d_status, nchunks * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_decompress")) != MSCOMP_OK) return st;
        st = lznt1_replay(nchunks, cap, tail, [&](size_t c) { return (size_t)sizes[c]; }, [&](size_t c) { return status[c]; }, &total);
        if (st == MSCOMP_OK && r.short_units) {
            uint8_t* d_packed = (uint8_t*)ctx->get(SC_AUX0, (size_t)r.bytes + 64);
            uint64_t* d_dst = (uint64_t*)ctx->get(SC_AUX1, (nchunks + 1) * sizeof(uint64_t));
            if (!d_packed || !d_dst) return MSCOMP_MEM_ERROR;
            if ((st = msc_compact_slots(ctx, d_out, 4096, d_sizes, (uint32_t)nchunks, d_dst, d_packed)) != MSCOMP_OK) return st;
            d_final = d_packed;
        }
    }
    if (st != MSCOMP_OK) return st;
    if (total) cudaMemcpyAsync(out, d_final, total, cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_decompress")) != MSCOMP_OK) return st;
    *out_len = total;
    return MSCOMP_OK;
}

// =================================================================================================================
// Xpress-Huffman and plain Xpress: the unit is the buffer
// =================================================================================================================
namespace {

struct UnitBatch {
    mscomp_gpu_batch b;
    uint32_t* d_out_len;
    int32_t* d_status;
};

// Uploads `count` unit descriptors into the context's scratch and fills a device-side batch.
bool upload_units(mscomp_gpu_ctx* ctx, const std::vector<uint64_t>& in_off, const std::vector<uint32_t>& in_len,
                  const std::vector<uint64_t>& out_off, const std::vector<uint32_t>& out_cap, UnitBatch* ub)
{
    const size_t n = in_off.size();
    const size_t bytes = n * (8 + 8 + 4 + 4 + 4 + 4);
    uint8_t* d = (uint8_t*)ctx->get(SC_AUX3, bytes + 64);
    uint8_t* h = (uint8_t*)ctx->get_pinned_big(bytes + 64);
    if (!d || !h) return false;
    memcpy(h, in_off.data(), n * 8);
    memcpy(h + n * 8, out_off.data(), n * 8);
    memcpy(h + n * 16, in_len.data(), n * 4);
    memcpy(h + n * 20, out_cap.data(), n * 4);
    memset(h + n * 24, 0, n * 8);
    cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream);
    ub->b.count = (uint32_t)n;
    ub->b.d_in_off = (const uint64_t*)d;
    ub->b.d_out_off = (const uint64_t*)(d + n * 8);
    ub->b.d_in_len = (const uint32_t*)(d + n * 16);
    ub->b.d_out_cap = (const uint32_t*)(d + n * 20);
    ub->b.d_out_len = ub->d_out_len = (uint32_t*)(d + n * 24);
    ub->b.d_status = ub->d_status = (int32_t*)(d + n * 28);
    return true;
}

typedef MSCompStatus (*batch_fn)(mscomp_gpu_ctx*, const uint8_t*, uint8_t*, const mscomp_gpu_batch*);

// One buffer through a batch entry point: host in -> device -> host out.  dev_cap = capacity given to the kernel.
MSCompStatus one_buffer(batch_fn fn, const char* what, const_bytes in, size_t in_len, bytes out, size_t* out_len, size_t dev_cap)
{
    if (!out_len || (!in && in_len) || (!out && *out_len)) return MSCOMP_ARG_ERROR;
    if (in_len > 0xFFFFFF00ull) return MSCOMP_ARG_ERROR;          // the device batch interface uses 32-bit lengths (with headroom
                                                                  // for the readers' rounding of end positions)
    const size_t cap = *out_len;
    if (dev_cap > 0xFFFFFF00ull) dev_cap = 0xFFFFFF00ull;
    mscomp_gpu_ctx* ctx = msc_thread_ctx();
    if (!ctx) return MSCOMP_MEM_ERROR;
    uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + 64);
    uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, dev_cap + 64);
    struct { uint32_t len; int32_t status; }* h_res = (decltype(h_res))ctx->get_pinned(256);
    if (!d_in || !d_out || !h_res) return MSCOMP_MEM_ERROR;
    UnitBatch ub;
    if (!upload_units(ctx, {0}, {(uint32_t)in_len}, {0}, {(uint32_t)dev_cap}, &ub)) return MSCOMP_MEM_ERROR;
    if (in_len) cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
    MSCompStatus st = fn(ctx, d_in, d_out, &ub.b);
    if (st != MSCOMP_OK) return st;
    cudaMemcpyAsync(&h_res->len, ub.d_out_len, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&h_res->status, ub.d_status, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), what)) != MSCOMP_OK) return st;
    if (h_res->status < 0) return (MSCompStatus)h_res->status;
    const size_t total = h_res->len;
    if (total > cap) return MSCOMP_BUF_ERROR;
    if (total) cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), what)) != MSCOMP_OK) return st;
    *out_len = total;
    return MSCOMP_OK;
}
} // namespace

extern "C" size_t xpress_max_compressed_size(size_t in_len) { return in_len + 4 + 4 * (in_len / 32); }
extern "C" size_t xpress_huff_max_compressed_size(size_t in_len) { return in_len + 34 + 258 + 258 * (in_len / 65536); }

extern "C" MSCompStatus xpress_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{   // one buffer = one stream = one thread block (the format has no independent sub-units)
    return one_buffer(mscomp_gpu_xpress_compress_batch, "xpress_compress", in, in_len, out, out_len, xpress_max_compressed_size(in_len));
}
extern "C" MSCompStatus xpress_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || (!out && *out_len)) return MSCOMP_ARG_ERROR;
    // One stream is a serial chain for one warp (the format has no marker to start a second reader from).  Streams of some size first
    // try the speculative path: the flag groups that could start at ANY offset are walked side by side, the true ones are linked up
    // and become a restart index (kernels/xpress_spec.cuh); whatever it cannot vouch for goes to the serial decoder below, which owns
    // the status codes.  MSCOMP_XPRESS_ONESHOT=serial forces that one.
    static const bool no_spec = getenv("MSCOMP_XPRESS_ONESHOT") && strcmp(getenv("MSCOMP_XPRESS_ONESHOT"), "serial") == 0;
    if (!no_spec && in_len >= 32768 && in_len <= (1u << 28) && *out_len) {
        mscomp_gpu_ctx* ctx = msc_thread_ctx();
        if (!ctx) return MSCOMP_MEM_ERROR;
        const size_t cap = *out_len > 0x7FFFFFFFull ? (size_t)0x7FFFFFFFull : *out_len;
        uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + 64);
        // (the room for the output is only taken once the stream's size is known: callers pass generous capacities)
        if (d_in) {
            cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
            uint64_t total = 0;
            bool handled = false;
            MSCompStatus st = msc_xpress_stream_decode(ctx, d_in, (uint32_t)in_len, nullptr, cap, &total, &handled);
            if (st != MSCOMP_OK) return st;
            if (handled) {
                uint8_t* d_out = (uint8_t*)ctx->scratch[SC_OUT];
                if (total) cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
                if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "xpress_decompress")) != MSCOMP_OK) return st;
                *out_len = total;
                return MSCOMP_OK;
            }
        }
    }
    return one_buffer(mscomp_gpu_xpress_decompress_batch, "xpress_decompress", in, in_len, out, out_len, out_len ? *out_len : 0);
}
extern "C" MSCompStatus xpress_huff_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || (!out && *out_len)) return MSCOMP_ARG_ERROR;
    // A multi-chunk stream is a serial chain for one lane (chunk lengths are implicit).  Streams of some size first try the
    // chunk-parallel path: candidate chunk starts are recognised by their complete code-length tables and parsed side by side
    // (kernels/xhuff_chunks.cuh); whatever it cannot vouch for goes to the serial decoder below, which owns the status codes.
    static const bool no_spec = getenv("MSCOMP_XHUFF_ONESHOT") && strcmp(getenv("MSCOMP_XHUFF_ONESHOT"), "serial") == 0;
    if (!no_spec && in_len >= 2048 && in_len <= 0xFFFFFF00ull) {             // (below that the serial decoder's one lane is as quick)
        mscomp_gpu_ctx* ctx = msc_thread_ctx();
        if (!ctx) return MSCOMP_MEM_ERROR;
        const size_t cap = *out_len > 0xFFFFFF00ull ? (size_t)0xFFFFFF00ull : *out_len;
        uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + 64);
        uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, cap + 64);
        if (d_in && d_out) {
            cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
            uint64_t total = 0;
            bool handled = false;
            MSCompStatus st = msc_xhuff_stream_decode(ctx, d_in, (uint32_t)in_len, d_out, cap, &total, &handled);
            if (st != MSCOMP_OK) return st;
            if (handled) {
                if (total) cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
                if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "xpress_huff_decompress")) != MSCOMP_OK) return st;
                *out_len = total;
                return MSCOMP_OK;
            }
        }
    }
    return one_buffer(mscomp_gpu_xpress_huff_decompress_batch, "xpress_huff_decompress", in, in_len, out, out_len, out_len ? *out_len : 0);
}

extern "C" MSCompStatus xpress_huff_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || (!out && *out_len)) return MSCOMP_ARG_ERROR;
    if (in_len == 0) { *out_len = 0; return MSCOMP_OK; }                     // src/xpress_huff_compress.cpp:249
    if (in_len > 0xFFFFFFFF0000ull) return MSCOMP_ARG_ERROR;
    const size_t cap = *out_len;
    mscomp_gpu_ctx* ctx = msc_thread_ctx();
    if (!ctx) return MSCOMP_MEM_ERROR;
    // Every 64 KiB chunk is encoded by its own thread block into a slot (eos_mode 1: the units are the chunks of ONE stream, so the
    // reference-parse encoder sees the reference's tokens -- window 65,535 across chunk starts; the fast matcher resets the window per
    // chunk, which the reference's decoder accepts as well: SURVEY.md 0.2); the slots are then packed into one stream.
    const size_t nunits = (in_len + 65535) / 65536;
    const size_t stride = 65536 + 256 + 64;
    uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + 64);
    uint8_t* d_slots = (uint8_t*)ctx->get(SC_SLOTS, nunits * stride + 64);
    uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, nunits * stride + 64);
    uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (nunits + 1) * sizeof(uint64_t));
    if (!d_in || !d_slots || !d_out || !d_offs) return MSCOMP_MEM_ERROR;
    std::vector<uint64_t> in_off(nunits), out_off(nunits);
    std::vector<uint32_t> in_len_u(nunits), out_cap(nunits, (uint32_t)stride);
    for (size_t u = 0; u < nunits; ++u) {
        in_off[u] = u * 65536; out_off[u] = u * stride;
        in_len_u[u] = (uint32_t)(in_len - u * 65536 < 65536 ? in_len - u * 65536 : 65536);
    }
    UnitBatch ub;
    if (!upload_units(ctx, in_off, in_len_u, out_off, out_cap, &ub)) return MSCOMP_MEM_ERROR;
    struct Res { uint64_t total; int32_t status; uint32_t pad; };
    mscomp_gpu_result* d_res = (mscomp_gpu_result*)ctx->get(SC_RESULT, 256);
    mscomp_gpu_result* h_res = (mscomp_gpu_result*)ctx->get_pinned(256);
    if (!d_res || !h_res) return MSCOMP_MEM_ERROR;
    cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
    MSCompStatus st = msc_xpress_huff_compress_units(ctx, d_in, d_slots, &ub.b, 1);
    if (st != MSCOMP_OK) return st;
    if ((st = msc_summarize(ctx, ub.d_out_len, ub.d_status, (uint32_t)nunits, d_res)) != MSCOMP_OK) return st;
    if ((st = msc_compact_slots(ctx, d_slots, stride, ub.d_out_len, (uint32_t)nunits, d_offs, d_out)) != MSCOMP_OK) return st;
    cudaMemcpyAsync(h_res, d_res, sizeof(*h_res), cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "xpress_huff_compress")) != MSCOMP_OK) return st;
    if (h_res->status < 0) return (MSCompStatus)h_res->status;
    const size_t total = (size_t)h_res->bytes;
    if (total > cap) return MSCOMP_BUF_ERROR;
    cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream);
    if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "xpress_huff_compress")) != MSCOMP_OK) return st;
    *out_len = total;
    return MSCOMP_OK;
}

// =================================================================================================================
// format dispatch (src/mscomp.cpp:87-132)
// =================================================================================================================
static MSCompStatus copy_engine(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{   // MSCOMP_NONE: src/mscomp.cpp:26-32
    if (!out_len) return MSCOMP_ARG_ERROR;
    if (in_len > *out_len) return MSCOMP_BUF_ERROR;
    if (in_len) memcpy(out, in, in_len);
    *out_len = in_len;
    return MSCOMP_OK;
}
static size_t copy_max_size(size_t in_len) { return in_len; }

typedef MSCompStatus (*codec_fn)(const_bytes, size_t, bytes, size_t*);
typedef size_t (*size_fn)(size_t);
static const codec_fn kCompress[5]   = { copy_engine, nullptr, lznt1_compress,   MSC_XPRESS_COMPRESS,   MSC_XPRESS_HUFF_COMPRESS };
static const codec_fn kDecompress[5] = { copy_engine, nullptr, lznt1_decompress, MSC_XPRESS_DECOMPRESS, MSC_XPRESS_HUFF_DECOMPRESS };
static const size_fn  kMaxSize[5]    = { copy_max_size, nullptr, lznt1_max_compressed_size, MSC_XPRESS_MAX, MSC_XPRESS_HUFF_MAX };

extern "C" MSCompStatus ms_compress(MSCompFormat format, const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if ((unsigned)format >= 5 || !kCompress[format]) return MSCOMP_ARG_ERROR;
    return kCompress[format](in, in_len, out, out_len);
}
extern "C" MSCompStatus ms_decompress(MSCompFormat format, const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if ((unsigned)format >= 5 || !kDecompress[format]) return MSCOMP_ARG_ERROR;
    return kDecompress[format](in, in_len, out, out_len);
}
extern "C" size_t ms_max_compressed_size(MSCompFormat format, size_t in_len)
{
    if ((unsigned)format >= 5 || !kMaxSize[format]) return (size_t)-1;
    return kMaxSize[format](in_len);
}
