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
#include <string.h>
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

// =================================================================================================================
// LZNT1
// =================================================================================================================
extern "C" size_t lznt1_max_compressed_size(size_t in_len) { return in_len + 3 + 2 * ((in_len + 4095) / 4096); }

extern "C" MSCompStatus lznt1_compress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || !out) return MSCOMP_ARG_ERROR;
    const size_t cap = *out_len;
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
    const size_t kSlice = (size_t)64 << 20;                 // upload / encode granularity (a multiple of the chunk size)
    if (in_len > kSlice && ctx->copy_ready()) {
        // chunks are independent: upload slice i+1 on the copy stream while slice i is being encoded
        uint8_t* d_slots = (uint8_t*)ctx->get(SC_SLOTS, nchunks * MSC_LZNT1_SLOT_BYTES + 16);
        uint32_t* d_sizes = (uint32_t*)ctx->get(SC_SIZES, (nchunks + 1) * sizeof(uint32_t));
        uint64_t* d_offs = (uint64_t*)ctx->get(SC_OFFS, (nchunks + 1) * sizeof(uint64_t));
        if (!d_slots || !d_sizes || !d_offs) return MSCOMP_MEM_ERROR;
        cudaEventRecord(ctx->copy_ev[15], ctx->stream);                      // earlier work on the scratch buffers
        cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[15], 0);
        int k = 0;
        for (size_t off = 0; off < in_len; off += kSlice, ++k) {
            const size_t len = in_len - off < kSlice ? in_len - off : kSlice;
            cudaEvent_t ev = ctx->copy_ev[k % 15];
            if (k >= 15) cudaEventSynchronize(ev);                            // (ring reuse: that slice was consumed long ago)
            cudaMemcpyAsync(d_in + off, in + off, len, cudaMemcpyHostToDevice, ctx->copy_stream);
            cudaEventRecord(ev, ctx->copy_stream);
            cudaStreamWaitEvent(ctx->stream, ev, 0);
            if ((st = msc_lznt1_encode_chunks(ctx, d_in, in_len, (uint32_t)(off / 4096), (uint32_t)((len + 4095) / 4096), d_slots, d_sizes)) != MSCOMP_OK) return st;
        }
        if ((st = msc_compact_slots(ctx, d_slots, MSC_LZNT1_SLOT_BYTES, d_sizes, (uint32_t)nchunks, d_offs, d_out)) != MSCOMP_OK) return st;
        cudaMemsetAsync(d_res, 0, sizeof(*d_res), ctx->stream);
        cudaMemcpyAsync(&d_res->bytes, d_offs + nchunks, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream);
    } else {
        if (in_len) cudaMemcpyAsync(d_in, in, in_len, cudaMemcpyHostToDevice, ctx->stream);
        if ((st = mscomp_gpu_lznt1_compress(ctx, d_in, in_len, d_out, dcap, nullptr, d_res)) != MSCOMP_OK) return st;
    }
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

namespace {
enum Lznt1Tail { T_CLEAN, T_ZERO, T_GARBAGE, T_TERM_OK, T_TERM_BAD, T_TRUNC, T_BADSIG };

// Replays the reference's sequential driver (src/lznt1_decompress.cpp:254-273 + internal.h:616-630) over the
// per-chunk results to obtain its status code.  size_of / status_of give chunk c's decoded size and status.
template <class FS, class FT>
MSCompStatus lznt1_replay(size_t nchunks, size_t cap, Lznt1Tail tail, FS size_of, FT status_of, size_t* total)
{
    size_t op = 0;
    for (size_t c = 0; c < nchunks; ++c) {
        if (cap - op == 0) return MSCOMP_BUF_ERROR;        // output full while input remains
        if (status_of(c) < 0) return MSCOMP_DATA_ERROR;
        const size_t sz = size_of(c);
        if (sz > cap - op) return MSCOMP_BUF_ERROR;        // decoded bytes left over in the bounce buffer
        op += sz;
    }
    *total = op;
    switch (tail) {
    case T_CLEAN: case T_ZERO: return MSCOMP_OK;
    case T_GARBAGE: return MSCOMP_BUF_ERROR;
    default: break;
    }
    if (cap - op == 0) return MSCOMP_BUF_ERROR;            // the loop never looks at the next header (:254)
    switch (tail) {
    case T_TERM_OK: return MSCOMP_OK;
    case T_TRUNC: return MSCOMP_BUF_ERROR;
    default: return MSCOMP_DATA_ERROR;                     // T_TERM_BAD, T_BADSIG
    }
}
} // namespace

extern "C" MSCompStatus lznt1_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{
    if (!out_len || (!in && in_len) || (!out && *out_len)) return MSCOMP_ARG_ERROR;
    const size_t cap = *out_len;

    // The upload of the stream starts first, on the copy stream, so that it overlaps with the header walk below.
    mscomp_gpu_ctx* ctx = in_len >= 2 ? msc_thread_ctx() : nullptr;
    if (in_len >= 2 && !ctx) return MSCOMP_MEM_ERROR;
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
    Lznt1Tail tail = T_CLEAN;
    size_t ip = 0;
    bool stopped = false;
    while (in_len - ip >= 2) {
        const unsigned h = rd16(in + ip);
        if (h == 0) { tail = (in_len - ip == 2) ? T_TERM_OK : T_TERM_BAD; stopped = true; break; }
        const size_t sz = (h & 0x0FFF) + 3;
        if (sz > in_len - ip) { tail = T_TRUNC; stopped = true; break; }
        if ((h & 0x7000) != 0x3000) { tail = T_BADSIG; stopped = true; break; }
        offs.push_back(ip);
        ip += sz;
    }
    if (!stopped) tail = (in_len - ip == 0) ? T_CLEAN : (in[ip] == 0 ? T_ZERO : T_GARBAGE);
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
            msc_compact_slots(ctx, d_out, 4096, d_sizes, (uint32_t)nchunks, d_dst, d_packed);
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
    uint8_t* h = (uint8_t*)ctx->get_pinned(bytes + 512);
    if (!d || !h) return false;
    h += 256;                                         // the first 256 bytes of the mailbox are used for results
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
    if (in_len > 0xFFFFFFFFull) return MSCOMP_ARG_ERROR;          // the device batch interface uses 32-bit lengths
    const size_t cap = *out_len;
    if (dev_cap > 0xFFFFFFFFull) dev_cap = 0xFFFFFFFFull;
    mscomp_gpu_ctx* ctx = msc_thread_ctx();
    if (!ctx) return MSCOMP_MEM_ERROR;
    uint8_t* d_in = (uint8_t*)ctx->get(SC_IN, in_len + 64);
    uint8_t* d_out = (uint8_t*)ctx->get(SC_OUT, dev_cap + 64);
    struct { uint32_t len; int32_t status; }* h_res = (decltype(h_res))ctx->get_pinned(256);
    if (!d_in || !d_out || !h_res) return MSCOMP_MEM_ERROR;
    UnitBatch ub;
    if (!upload_units(ctx, {0}, {(uint32_t)in_len}, {0}, {(uint32_t)dev_cap}, &ub)) return MSCOMP_MEM_ERROR;
    h_res = (decltype(h_res))ctx->get_pinned(256);                // (the mailbox may have been reallocated)
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
    return one_buffer(mscomp_gpu_xpress_decompress_batch, "xpress_decompress", in, in_len, out, out_len, out_len ? *out_len : 0);
}
extern "C" MSCompStatus xpress_huff_decompress(const_bytes in, size_t in_len, bytes out, size_t* out_len)
{   // a multi-chunk stream is a serial chain (chunk lengths are implicit, matches cross chunks): one lane walks it
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
    // Every 64 KiB chunk is compressed on its own (window reset per chunk, valid for the reference's decoder:
    // SURVEY.md 0.2) by its own thread block into a slot; the slots are then packed into one stream.
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
