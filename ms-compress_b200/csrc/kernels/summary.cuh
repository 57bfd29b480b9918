// summary.cuh -- reduce per-unit sizes / statuses into one mscomp_gpu_result (single block).
#pragma once
#include "common.cuh"

namespace msc {

struct GpuResult {      // layout == mscomp_gpu_result (include/mscomp_gpu.h)
    uint64_t bytes; int32_t status; uint32_t first_bad; uint32_t short_units; uint32_t reserved;
};

constexpr int SUM_THREADS = 1024;

// full_unit: expected size of every unit but the last (4096 for LZNT1 chunks, 0 = do not count short units)
__global__ void __launch_bounds__(SUM_THREADS)
summarize_kernel(const uint32_t* __restrict__ sizes, const int32_t* __restrict__ status, uint32_t n, uint32_t full_unit,
                 GpuResult* __restrict__ res)
{
    __shared__ uint64_t s_bytes[32];
    __shared__ uint32_t s_bad[32], s_short[32];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t bytes = 0; uint32_t bad = 0xffffffffu, nshort = 0;
    for (uint32_t i = tid; i < n; i += SUM_THREADS) {
        bytes += sizes[i];
        if (status[i] < 0 && i < bad) bad = i;
        if (full_unit && i + 1 < n && sizes[i] != full_unit) ++nshort;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        bytes += __shfl_xor_sync(FULL, bytes, d);
        bad = min(bad, __shfl_xor_sync(FULL, bad, d));
        nshort += __shfl_xor_sync(FULL, nshort, d);
    }
    if (lane == 0) { s_bytes[warp] = bytes; s_bad[warp] = bad; s_short[warp] = nshort; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < SUM_THREADS / 32; ++w) { bytes += s_bytes[w]; bad = min(bad, s_bad[w]); nshort += s_short[w]; }
        res->bytes = bytes; res->status = (bad == 0xffffffffu) ? 0 : status[bad];
        res->first_bad = (bad == 0xffffffffu) ? 0 : bad; res->short_units = nshort; res->reserved = 0;
    }
}

// max over units of (off[i] + cap[i]): the extent of the output region a batch can touch (single block)
__global__ void __launch_bounds__(SUM_THREADS)
batch_extent_kernel(const uint64_t* __restrict__ off, const uint32_t* __restrict__ cap, uint32_t n, uint64_t* __restrict__ extent)
{
    __shared__ uint64_t s_max[32];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t m = 0;
    for (uint32_t i = tid; i < n; i += SUM_THREADS) { const uint64_t e = off[i] + cap[i]; if (e > m) m = e; }
#pragma unroll
    for (int d = 16; d; d >>= 1) { const uint64_t o = __shfl_xor_sync(FULL, m, d); if (o > m) m = o; }
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    if (tid == 0) { for (int w = 1; w < SUM_THREADS / 32; ++w) if (s_max[w] > m) m = s_max[w]; *extent = m; }
}

} // namespace msc
