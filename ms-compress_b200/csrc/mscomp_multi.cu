// mscomp_multi.cu -- the multi-GPU entry points of the C ABI (include/mscomp_gpu.h, "several GPUs").
//
// The units of all three formats are independent, so a job spread over N GPUs is N single-GPU jobs on
// contiguous unit ranges plus ONE exchange: every rank learns every rank's output sizes, so that it can place (or report) its
// part of the global stream.  That exchange is an ncclAllGather on the context's stream, behind this library's own calls --
// a C program reaches more than one GPU without Python (one process or thread per GPU, one context each).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a single-GPU user of the library needs no NCCL at all.
#include <dlfcn.h>
#include <string.h>
#include "ctx.cuh"
#include "shim_common.h"

namespace {
// the handful of NCCL entry points used (signatures of nccl.h 2.x; ncclUniqueId is 128 opaque bytes, ncclUint64 = 5)
struct NcclId { char internal[128]; };
typedef int (*GetUniqueId_t)(NcclId*);
typedef int (*CommInitRank_t)(void** comm, int nranks, NcclId id, int rank);
typedef int (*CommDestroy_t)(void* comm);
typedef int (*AllGather_t)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t stream);
typedef const char* (*GetErrorString_t)(int);
struct Nccl {
    void* so = nullptr;
    GetUniqueId_t GetUniqueId = nullptr;
    CommInitRank_t CommInitRank = nullptr;
    CommDestroy_t CommDestroy = nullptr;
    AllGather_t AllGather = nullptr;
    GetErrorString_t GetErrorString = nullptr;
    bool ok = false;
};
Nccl& nccl()
{
    static Nccl n = [] {
        Nccl x;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { x.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (x.so) break; }
        if (!x.so) { fprintf(stderr, "libMSCompression(B200): NCCL not found (%s)\n", dlerror()); return x; }
        x.GetUniqueId = (GetUniqueId_t)dlsym(x.so, "ncclGetUniqueId");
        x.CommInitRank = (CommInitRank_t)dlsym(x.so, "ncclCommInitRank");
        x.CommDestroy = (CommDestroy_t)dlsym(x.so, "ncclCommDestroy");
        x.AllGather = (AllGather_t)dlsym(x.so, "ncclAllGather");
        x.GetErrorString = (GetErrorString_t)dlsym(x.so, "ncclGetErrorString");
        x.ok = x.GetUniqueId && x.CommInitRank && x.CommDestroy && x.AllGather;
        return x;
    }();
    return n;
}
MSCompStatus nccl_status(int rc, const char* what)
{
    if (rc == 0) return MSCOMP_OK;
    fprintf(stderr, "libMSCompression(B200): NCCL failure in %s: %s\n", what, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
    return MSCOMP_ERRNO;
}
} // namespace

// Contiguous unit range of a rank: [floor(rank * units / nranks), floor((rank + 1) * units / nranks)).
extern "C" void mscomp_gpu_shard_range(uint64_t units, uint32_t rank, uint32_t nranks, uint64_t* first, uint64_t* count)
{
    if (!nranks) nranks = 1;
    if (rank >= nranks) rank = nranks - 1;
    const unsigned __int128 u = units;
    const uint64_t lo = (uint64_t)(u * rank / nranks), hi = (uint64_t)(u * (rank + 1) / nranks);
    if (first) *first = lo;
    if (count) *count = hi - lo;
}

extern "C" MSCompStatus mscomp_gpu_comm_unique_id(void* id)
{
    if (!id) return MSCOMP_ARG_ERROR;
    if (!nccl().ok) return MSCOMP_MEM_ERROR;
    NcclId x;
    const MSCompStatus st = nccl_status(nccl().GetUniqueId(&x), "ncclGetUniqueId");
    if (st == MSCOMP_OK) memcpy(id, &x, sizeof(x));
    return st;
}

extern "C" MSCompStatus mscomp_gpu_comm_init(mscomp_gpu_ctx* ctx, const void* id, uint32_t rank, uint32_t nranks)
{
    if (!ctx || !id || !nranks || rank >= nranks || ctx->comm) return MSCOMP_ARG_ERROR;
    if (!nccl().ok) return MSCOMP_MEM_ERROR;
    cudaSetDevice(ctx->device);
    NcclId x;
    memcpy(&x, id, sizeof(x));
    void* comm = nullptr;
    const MSCompStatus st = nccl_status(nccl().CommInitRank(&comm, (int)nranks, x, (int)rank), "ncclCommInitRank");
    if (st != MSCOMP_OK) return st;
    ctx->comm = comm; ctx->comm_owned = true; ctx->comm_rank = rank; ctx->comm_size = nranks;
    return MSCOMP_OK;
}

extern "C" MSCompStatus mscomp_gpu_comm_attach(mscomp_gpu_ctx* ctx, void* nccl_comm, uint32_t rank, uint32_t nranks)
{
    if (!ctx || !nccl_comm || !nranks || rank >= nranks || ctx->comm) return MSCOMP_ARG_ERROR;
    if (!nccl().ok) return MSCOMP_MEM_ERROR;
    ctx->comm = nccl_comm; ctx->comm_owned = false; ctx->comm_rank = rank; ctx->comm_size = nranks;
    return MSCOMP_OK;
}

extern "C" void mscomp_gpu_comm_release(mscomp_gpu_ctx* ctx)
{
    if (!ctx || !ctx->comm) return;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_owned && nccl().ok) nccl().CommDestroy(ctx->comm);
    ctx->comm = nullptr; ctx->comm_owned = false; ctx->comm_rank = 0; ctx->comm_size = 1;
}

extern "C" uint32_t mscomp_gpu_comm_rank(const mscomp_gpu_ctx* ctx) { return ctx ? ctx->comm_rank : 0; }
extern "C" uint32_t mscomp_gpu_comm_size(const mscomp_gpu_ctx* ctx) { return ctx ? ctx->comm_size : 1; }

// d_all[r * n_values + k] = rank r's d_mine[k], on the context's stream.  Without a communicator (one GPU) it is a copy.
extern "C" MSCompStatus mscomp_gpu_sizes_allgather(mscomp_gpu_ctx* ctx, const uint64_t* d_mine, uint32_t n_values, uint64_t* d_all)
{
    if (!ctx || !d_mine || !d_all || !n_values) return MSCOMP_ARG_ERROR;
    cudaSetDevice(ctx->device);
    if (!ctx->comm) return msc_cuda_status(cudaMemcpyAsync(d_all, d_mine, (size_t)n_values * 8, cudaMemcpyDeviceToDevice, ctx->stream), "sizes_allgather");
    ++ctx->collectives;
    return nccl_status(nccl().AllGather(d_mine, d_all, n_values, /*ncclUint64*/ 5, ctx->comm, ctx->stream), "ncclAllGather");
}
extern "C" uint64_t mscomp_gpu_collectives(const mscomp_gpu_ctx* ctx) { return ctx ? ctx->collectives : 0; }

// This rank's shard of an LZNT1 stream (a whole number of 4 KiB chunks, except on the last rank) -> its part of the compressed
// stream, and every rank's compressed size: d_sizes_all[r] = bytes of rank r; the global stream is the parts in rank order.
extern "C" MSCompStatus mscomp_gpu_lznt1_compress_sharded(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                                                          uint64_t* d_chunk_off, mscomp_gpu_result* d_result, uint64_t* d_sizes_all)
{
    if (!ctx || !d_result || !d_sizes_all) return MSCOMP_ARG_ERROR;
    const MSCompStatus st = mscomp_gpu_lznt1_compress(ctx, d_in, in_len, d_out, out_cap, d_chunk_off, d_result);
    if (st != MSCOMP_OK) return st;
    return mscomp_gpu_sizes_allgather(ctx, &d_result->bytes, 1, d_sizes_all);      // (bytes is the first member of the result)
}

// A batch of independent buffers (Xpress-Huffman: format 4, plain Xpress: format 3), this rank's range of it: compress, then
// d_sizes_all[r] = total compressed bytes of rank r (d_my_total: one device uint64 of scratch for this rank's sum).
namespace {
__global__ void sum_lengths_kernel(const uint32_t* __restrict__ len, const int32_t* __restrict__ status, uint32_t n, unsigned long long* __restrict__ out)
{
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (status[i] == 0) ? len[i] : 0u;
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}
} // namespace
extern "C" MSCompStatus mscomp_gpu_compress_batch_sharded(mscomp_gpu_ctx* ctx, int format, const uint8_t* d_in, uint8_t* d_out, const mscomp_gpu_batch* batch,
                                                          uint64_t* d_my_total, uint64_t* d_sizes_all)
{
    if (!ctx || !batch || !d_my_total || !d_sizes_all) return MSCOMP_ARG_ERROR;
    MSCompStatus st;
    if (format == MSCOMP_XPRESS_HUFF) st = mscomp_gpu_xpress_huff_compress_batch(ctx, d_in, d_out, batch);
    else if (format == MSCOMP_XPRESS) st = mscomp_gpu_xpress_compress_batch(ctx, d_in, d_out, batch);
    else return MSCOMP_ARG_ERROR;
    if (st != MSCOMP_OK) return st;
    cudaMemsetAsync(d_my_total, 0, 8, ctx->stream);
    if (batch->count) MSC_LAUNCH(ctx, sum_lengths_kernel, 64, 256, (const uint32_t*)batch->d_out_len, (const int32_t*)batch->d_status, batch->count, (unsigned long long*)d_my_total);
    return mscomp_gpu_sizes_allgather(ctx, d_my_total, 1, d_sizes_all);
}
