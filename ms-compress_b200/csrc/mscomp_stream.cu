// mscomp_stream.cu -- the reference's streaming API for LZNT1 (and the MSCOMP_NONE copy stream): the state machine of
// stream_logic.h with the chunk codecs running on the GPU -- a call gathers all the chunks it can process and hands them to
// the device API as ONE batch.  No codec work happens on the CPU.
#include "ctx.cuh"
#include <new>
#include <stdlib.h>
#include "shim_common.h"
#include "stream_logic.h"
#include "kernels/xpress_stream.cuh"
#include "../../include/xpress.h"

using namespace msc_stream;

namespace {
struct GpuCodec {
    // ---- GPU batches ------------------------------------------------------------------------------------------------------
    // Encodes in[0..n) as LZNT1 chunks (the last one may be short): comp = the chunks back to back (headers included),
    // offs[c] = start of chunk c in comp, offs[nchunks] = total.
    static MSCompStatus encode(const uint8_t* in, size_t n, std::vector<uint8_t>& comp, std::vector<uint64_t>& offs)
    {
        mscomp_gpu_ctx* ctx = msc_thread_ctx();
        if (!ctx) return MSCOMP_MEM_ERROR;
        const size_t nch = (n + CHUNK - 1) / CHUNK, cap = n + 2 * nch + 16;
        uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, n + 64);
        uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, cap + 64);
        uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (nch + 1) * sizeof(uint64_t));
        mscomp_gpu_result* d_res = (mscomp_gpu_result*)ctx->get(SC_RESULT, 256);
        if (!d_in || !d_out || !d_offs || !d_res) return MSCOMP_MEM_ERROR;
        cudaMemcpyAsync(d_in, in, n, cudaMemcpyHostToDevice, ctx->stream);
        MSCompStatus st = mscomp_gpu_lznt1_compress(ctx, d_in, n, d_out, cap, d_offs, d_res);
        if (st != MSCOMP_OK) return st;
        offs.resize(nch + 1);
        cudaMemcpyAsync(offs.data(), d_offs, (nch + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if ((st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_deflate")) != MSCOMP_OK) return st;
        comp.resize((size_t)offs[nch]);
        if (!comp.empty()) cudaMemcpyAsync(comp.data(), d_out, comp.size(), cudaMemcpyDeviceToHost, ctx->stream);
        return msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_deflate");
    }

    // Decodes the compressed chunks whose headers sit at in[offs[i]]: chunk i -> out[4096 * i ..], sizes[i], status[i].
    static MSCompStatus decode(const uint8_t* in, size_t n, const std::vector<uint64_t>& offs, std::vector<uint8_t>& out,
                            std::vector<uint32_t>& sizes, std::vector<int32_t>& status)
    {
        mscomp_gpu_ctx* ctx = msc_thread_ctx();
        if (!ctx) return MSCOMP_MEM_ERROR;
        const size_t k = offs.size();
        uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, n + 64);
        uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, k * CHUNK + 64);
        uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (k + 1) * sizeof(uint64_t));
        uint32_t* d_sizes = (uint32_t*)ctx->get(SC_SIZES, (k + 1) * sizeof(uint32_t));
        int32_t* d_status = (int32_t*)ctx->get(SC_STATUS, (k + 1) * sizeof(int32_t));
        mscomp_gpu_result* d_res = (mscomp_gpu_result*)ctx->get(SC_RESULT, 256);
        if (!d_in || !d_out || !d_offs || !d_sizes || !d_status || !d_res) return MSCOMP_MEM_ERROR;
        cudaMemcpyAsync(d_in, in, n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_offs, offs.data(), k * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        MSCompStatus st = mscomp_gpu_lznt1_decompress(ctx, d_in, n, d_offs, (uint32_t)k, d_out, d_sizes, d_status, d_res);
        if (st != MSCOMP_OK) return st;
        out.resize(k * CHUNK); sizes.resize(k); status.resize(k);
        cudaMemcpyAsync(out.data(), d_out, k * CHUNK, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(sizes.data(), d_sizes, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(status.data(), d_status, k * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        return msc_cuda_status(cudaStreamSynchronize(ctx->stream), "lznt1_inflate");
    }

};
} // namespace

namespace {
// xpress_inflate on the GPU: the decoder state (8 KiB history included) stays in device memory between calls; a call uploads the
// input it was given, runs one step (one thread: the stream is byte-serial), and downloads the bytes produced.
struct GpuXpressStream {
    struct Handle { msc::XpStreamState* d_state; };
    static void* create()
    {
        mscomp_gpu_ctx* ctx = msc_thread_ctx();
        if (!ctx) return nullptr;
        Handle* h = new (std::nothrow) Handle{nullptr};
        if (!h) return nullptr;
        if (cudaMalloc((void**)&h->d_state, sizeof(msc::XpStreamState)) != cudaSuccess) { cudaGetLastError(); delete h; return nullptr; }
        static const bool quirk = getenv("MSCOMP_XPRESS_INFLATE_UPSTREAM_QUIRK") != nullptr;     // test hook, see kernels/xpress_stream.cuh (step A)
        MSC_LAUNCH(ctx, msc::xpress_stream_init_kernel, 1, 32, h->d_state, quirk ? 1u : 0u);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(h->d_state); delete h; return nullptr; }
        return h;
    }
    static void destroy(void* hv)
    {
        Handle* h = (Handle*)hv;
        if (mscomp_gpu_ctx* ctx = msc_thread_ctx()) cudaStreamSynchronize(ctx->stream);
        cudaFree(h->d_state);
        delete h;
    }
    static MSCompStatus step(void* hv, const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t n_out, uint32_t* used, uint32_t* made, int* status)
    {
        Handle* h = (Handle*)hv;
        mscomp_gpu_ctx* ctx = msc_thread_ctx();
        if (!ctx) return MSCOMP_MEM_ERROR;
        uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, (size_t)n_in + 64);
        uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, (size_t)n_out + 64);
        msc::XpStreamResult* d_res = (msc::XpStreamResult*)ctx->get(SC_RESULT, 256);
        msc::XpStreamResult* h_res = (msc::XpStreamResult*)ctx->get_pinned(256);
        if (!d_in || !d_out || !d_res || !h_res) return MSCOMP_MEM_ERROR;
        if (n_in) cudaMemcpyAsync(d_in, in, n_in, cudaMemcpyHostToDevice, ctx->stream);
        MSC_LAUNCH(ctx, msc::xpress_stream_step_kernel, 1, 32, h->d_state, (const uint8_t*)d_in, n_in, d_out, n_out, d_res);
        cudaMemcpyAsync(h_res, d_res, sizeof(*h_res), cudaMemcpyDeviceToHost, ctx->stream);
        MSCompStatus st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "xpress_inflate");
        if (st != MSCOMP_OK) return st;
        *used = h_res->in_used; *made = h_res->out_made; *status = h_res->status;
        if (*made) {
            cudaMemcpyAsync(out, d_out, *made, cudaMemcpyDeviceToHost, ctx->stream);
            st = msc_cuda_status(cudaStreamSynchronize(ctx->stream), "xpress_inflate");
        }
        return st;
    }
    static void tail(void* hv, uint32_t* flags, uint32_t* flagged, uint32_t* kept)
    {
        Handle* h = (Handle*)hv;
        uint32_t w[5] = {0, 0, 0, 0, 1};
        if (mscomp_gpu_ctx* ctx = msc_thread_ctx()) {
            cudaMemcpyAsync(w, h->d_state, sizeof(w), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
        }
        *flags = w[0]; *flagged = w[1]; *kept = w[4];
    }
};
} // namespace

extern "C" MSCompStatus xpress_deflate_init(mscomp_stream* s) { return xpress_deflate_init_impl(s); }
extern "C" MSCompStatus xpress_deflate(mscomp_stream* s, MSCompFlush flush) { return xpress_deflate_impl(s, flush); }
extern "C" MSCompStatus xpress_deflate_end(mscomp_stream* s) { return xpress_deflate_end_impl(s); }
extern "C" MSCompStatus xpress_inflate_init(mscomp_stream* s) { return xpress_inflate_init_impl<GpuXpressStream>(s); }
extern "C" MSCompStatus xpress_inflate(mscomp_stream* s) { return xpress_inflate_impl<GpuXpressStream>(s); }
extern "C" MSCompStatus xpress_inflate_end(mscomp_stream* s) { return xpress_inflate_end_impl<GpuXpressStream>(s); }

extern "C" MSCompStatus lznt1_deflate_init(mscomp_stream* s) { return lznt1_deflate_init_impl(s); }
extern "C" MSCompStatus lznt1_deflate(mscomp_stream* s, MSCompFlush flush) { return lznt1_deflate_impl<GpuCodec>(s, flush); }
extern "C" MSCompStatus lznt1_deflate_end(mscomp_stream* s) { return lznt1_deflate_end_impl(s); }
extern "C" MSCompStatus lznt1_inflate_init(mscomp_stream* s) { return lznt1_inflate_init_impl(s); }
extern "C" MSCompStatus lznt1_inflate(mscomp_stream* s) { return lznt1_inflate_impl<GpuCodec>(s); }
extern "C" MSCompStatus lznt1_inflate_end(mscomp_stream* s) { return lznt1_inflate_end_impl(s); }

// ---- format dispatch (src/mscomp.cpp:134-223) -------------------------------------------------------------------------
extern "C" MSCompStatus ms_deflate_init(MSCompFormat format, mscomp_stream* s)
{
    switch ((unsigned)format) {
    case MSCOMP_NONE: return copy_init(s);
    case MSCOMP_LZNT1: return lznt1_deflate_init(s);
    case MSCOMP_XPRESS: return xpress_deflate_init(s);
    default: set_error(s, "no streaming compressor for this format"); return MSCOMP_ARG_ERROR;
    }
}
extern "C" MSCompStatus ms_deflate(mscomp_stream* s, MSCompFlush flush)
{
    if (s && s->format == MSCOMP_NONE) return copy_deflate(s, flush);
    if (s && s->format == MSCOMP_LZNT1) return lznt1_deflate(s, flush);
    if (s && s->format == MSCOMP_XPRESS) return xpress_deflate(s, flush);
    set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size");
    return MSCOMP_ARG_ERROR;
}
extern "C" MSCompStatus ms_deflate_end(mscomp_stream* s)
{
    if (s && s->format == MSCOMP_NONE) return copy_end(s);
    if (s && s->format == MSCOMP_LZNT1) return lznt1_deflate_end(s);
    if (s && s->format == MSCOMP_XPRESS) return xpress_deflate_end(s);
    set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size");
    return MSCOMP_ARG_ERROR;
}
extern "C" MSCompStatus ms_inflate_init(MSCompFormat format, mscomp_stream* s)
{
    switch ((unsigned)format) {
    case MSCOMP_NONE: return copy_init(s);
    case MSCOMP_LZNT1: return lznt1_inflate_init(s);
    case MSCOMP_XPRESS: return xpress_inflate_init(s);
    default: set_error(s, "no streaming decompressor for this format"); return MSCOMP_ARG_ERROR;
    }
}
extern "C" MSCompStatus ms_inflate(mscomp_stream* s)
{
    if (s && s->format == MSCOMP_NONE) return copy_inflate(s);
    if (s && s->format == MSCOMP_LZNT1) return lznt1_inflate(s);
    if (s && s->format == MSCOMP_XPRESS) return xpress_inflate(s);
    set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size");
    return MSCOMP_ARG_ERROR;
}
extern "C" MSCompStatus ms_inflate_end(mscomp_stream* s)
{
    if (s && s->format == MSCOMP_NONE) return copy_end(s);
    if (s && s->format == MSCOMP_LZNT1) return lznt1_inflate_end(s);
    if (s && s->format == MSCOMP_XPRESS) return xpress_inflate_end(s);
    set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size");
    return MSCOMP_ARG_ERROR;
}
