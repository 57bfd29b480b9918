// compact.cuh -- size scan + slot compaction shared by the encoders.
//
// The chunk/buffer encoders write each unit into a fixed-stride slot and record its size; these two kernels
// turn that into the contiguous stream the reference's one-shot compressors produce (the reference simply
// advances an output pointer, e.g. src/lznt1_compress.cpp:262-263).
#pragma once
#include "common.cuh"

namespace msc {

constexpr int SCAN_THREADS = 1024;

// offs[i] = base + sum_{j<i} sizes[j] for i in [0, n]; single block (n is a unit count: <= a few million).
// base_ptr (may be null) adds a device-resident base: the end offset of the previous slice when a stream is compacted
// slice by slice.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_sizes_kernel(const uint32_t* __restrict__ sizes, uint32_t n, uint64_t base, const uint64_t* base_ptr, uint64_t* __restrict__ offs,
                  uint64_t* total_mirror)   // optional copy of offs[n] in mapped host memory
{
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = base + (base_ptr ? *base_ptr : 0);
    __syncthreads();
    for (uint32_t start = 0; start < n; start += SCAN_THREADS * 4) {
        uint32_t v[4];
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint32_t idx = start + tid * 4 + i; v[i] = idx < n ? sizes[idx] : 0; s += v[i]; }
        uint64_t incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint64_t woff = carry;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        uint64_t run = woff + incl - s;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint32_t idx = start + tid * 4 + i; if (idx < n) offs[idx] = run; run += v[i]; }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) carry = run;
        __syncthreads();
    }
    if (tid == 0) { offs[n] = carry; if (total_mirror) { *total_mirror = carry; __threadfence_system(); } }
}

// One warp per unit: dst[offs[u] .. +sizes[u]) = slots[u*stride .. +sizes[u])
__global__ void __launch_bounds__(256)
gather_slots_kernel(const uint8_t* __restrict__ slots, uint64_t stride, const uint32_t* __restrict__ sizes,
                    const uint64_t* __restrict__ offs, uint32_t n, uint8_t* __restrict__ dst)
{
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    for (uint32_t u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < n; u += warps_total)
        warp_copy_bytes(dst + offs[u], slots + (uint64_t)u * stride, sizes[u]);
}

} // namespace msc
