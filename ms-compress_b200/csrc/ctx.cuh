// ctx.cuh -- the GPU context behind mscomp_gpu_ctx: one CUDA stream + grow-only device/pinned scratch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mscomp_gpu.h"

enum ScratchId { SC_SLOTS = 0, SC_SIZES, SC_STATUS, SC_OFFS, SC_RESULT, SC_IN, SC_OUT, SC_AUX0, SC_AUX1, SC_AUX2, SC_AUX3, SC_AUX4, SC_COUNT };

struct mscomp_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    int encoder = 0;                 // MSCOMP_GPU_ENCODER_*: Xpress / Xpress-Huffman encoders (mscomp_gpu_set_encoder)
    void* scratch[SC_COUNT] = {};
    size_t scratch_cap[SC_COUNT] = {};
    cudaStream_t copy_stream = nullptr;   // upload stream of the host-pointer wrappers: uploads overlap with kernels
    cudaStream_t down_stream = nullptr;   // download stream: results leave while later slices are still being processed
    cudaEvent_t copy_ev[16] = {};         // upload events (ring)
    cudaEvent_t work_ev[16] = {};         // kernel-side events (ring)
    void* comm = nullptr;            // ncclComm_t of the multi-GPU calls (mscomp_multi.cu); null = one GPU
    bool comm_owned = false;
    uint32_t comm_rank = 0, comm_size = 1;
    uint64_t collectives = 0;        // NCCL collectives issued through this context
    void* pinned = nullptr;          // small pinned mailbox for results
    size_t pinned_cap = 0;

    // lazily created upload stream + a ring of events
    bool copy_ready()
    {
        if (copy_stream) return true;
        if (cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); copy_stream = nullptr; return false; }
        if (cudaStreamCreateWithFlags(&down_stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError(); cudaStreamDestroy(copy_stream); copy_stream = down_stream = nullptr; return false;
        }
        for (int i = 0; i < 16; ++i) { cudaEventCreateWithFlags(&copy_ev[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&work_ev[i], cudaEventDisableTiming); }
        return true;
    }

    // ---- optional per-kernel profiling ----
    static constexpr int kProfMax = 2048;
    bool prof_on = false;
    int prof_n = 0;
    cudaEvent_t* prof_ev = nullptr;          // 2 * kProfMax events, created on first use
    const char* prof_name[kProfMax];
    int prof_begin(const char* name)
    {
        if (!prof_on || prof_n >= kProfMax) return -1;
        if (!prof_ev) {
            prof_ev = new cudaEvent_t[2 * kProfMax];
            for (int i = 0; i < 2 * kProfMax; ++i) cudaEventCreate(&prof_ev[i]);
        }
        const int s = prof_n++;
        prof_name[s] = name;
        cudaEventRecord(prof_ev[2 * s], stream);
        return s;
    }
    void prof_end(int s) { if (s >= 0) cudaEventRecord(prof_ev[2 * s + 1], stream); }

    // grow-only device scratch; returns nullptr on allocation failure
    void* get(ScratchId id, size_t bytes)
    {
        if (id == SC_RESULT && bytes < 4096) bytes = 4096;   // the device result mailbox never moves: callers keep pointers into it across launches
        if (bytes <= scratch_cap[id]) return scratch[id];
        cudaStreamSynchronize(stream);              // the old block may still be in use
        if (scratch[id]) cudaFree(scratch[id]);
        scratch[id] = nullptr; scratch_cap[id] = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&scratch[id], want) != cudaSuccess) { cudaGetLastError(); scratch[id] = nullptr; return nullptr; }
        scratch_cap[id] = want;
        return scratch[id];
    }
    // The result mailbox: one fixed pinned block that is allocated once and never moves, so pointers into it stay valid
    // whatever else a call allocates (kernels write results straight into it through the mapped address).
    static constexpr size_t kMailboxBytes = 8192;
    void* get_pinned(size_t bytes)
    {
        if (bytes > kMailboxBytes) return nullptr;
        if (pinned) return pinned;
        if (cudaMallocHost(&pinned, kMailboxBytes) != cudaSuccess) { cudaGetLastError(); pinned = nullptr; return nullptr; }
        pinned_cap = kMailboxBytes;
        return pinned;
    }
    // Grow-only pinned staging for bulk host-side tables (unit descriptors); separate from the mailbox.
    void* pinned_big = nullptr;
    size_t pinned_big_cap = 0;
    void* get_pinned_big(size_t bytes)
    {
        if (bytes <= pinned_big_cap) return pinned_big;
        cudaStreamSynchronize(stream);              // an upload from the old block may still be in flight
        if (pinned_big) cudaFreeHost(pinned_big);
        pinned_big = nullptr; pinned_big_cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaMallocHost(&pinned_big, want) != cudaSuccess) { cudaGetLastError(); pinned_big = nullptr; return nullptr; }
        pinned_big_cap = want;
        return pinned_big;
    }
    // A few zero-initialised device counters (diagnostics), allocated on first use.
    unsigned long long* counters = nullptr;
    unsigned long long* get_counters()
    {
        if (counters) return counters;
        if (cudaMalloc((void**)&counters, 256) != cudaSuccess) { cudaGetLastError(); counters = nullptr; return nullptr; }
        cudaMemsetAsync(counters, 0, 256, stream);
        return counters;
    }
    // Kernels that need more than 48 KB of dynamic shared memory opt in once per context (the attribute is per device,
    // and a context is bound to one device and used by one host thread at a time).
    static constexpr int kOptInMax = 32;
    const void* opted_in[kOptInMax] = {};
    int opted_n = 0;
    void opt_in_smem(const void* func, size_t bytes)
    {
        for (int i = 0; i < opted_n; ++i) if (opted_in[i] == func) return;
        cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (opted_n < kOptInMax) opted_in[opted_n++] = func;
    }
};

// Optional per-kernel timing (mscomp_gpu_profile_*): CUDA events recorded on the context's stream around every
// launch; the benchmark uses it to report each kernel's device time for the roofline figures.
#define MSC_LAUNCH_SMEM(ctx, kernel, grid, block, smem, ...)                                        \
    do {                                                                                            \
        const int prof_slot_ = (ctx)->prof_begin(#kernel);                                          \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                            \
        ++(ctx)->launches;                                                                          \
        (ctx)->prof_end(prof_slot_);                                                                \
    } while (0)
#define MSC_LAUNCH(ctx, kernel, grid, block, ...) MSC_LAUNCH_SMEM(ctx, kernel, grid, block, 0, __VA_ARGS__)

// Kernels that need more than 48 KB of dynamic shared memory opt in once per context (= per device).
#define MSC_OPT_IN_SMEM(kernel, bytes) (ctx)->opt_in_smem((const void*)(kernel), (bytes))

static inline MSCompStatus msc_cuda_status(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return MSCOMP_OK;
    fprintf(stderr, "libMSCompression(B200): CUDA failure in %s: %s\n", what, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? MSCOMP_MEM_ERROR : MSCOMP_ERRNO;
}
