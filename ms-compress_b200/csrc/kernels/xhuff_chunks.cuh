// xhuff_chunks.cuh -- chunk-parallel parsing of ONE multi-chunk Xpress-Huffman stream (the one-shot xpress_huff_decompress).
//
// The format stores no chunk lengths (reference src/xpress_huff_decompress.cpp:136-162: a chunk ends when it has produced
// 64 KiB, and only then is the next chunk's start known), so a stream is a serial chain for one GPU lane.  What the stream DOES
// offer is a strong mark at every chunk start: 256 bytes of 4-bit code lengths of a COMPLETE prefix code -- their Kraft sum
// sum(2^(15-len)) is exactly 2^15, which a random 256-byte window of compressed data essentially never meets.  So:
//   1. xhs_candidates_kernel   every offset whose next 256 bytes are a complete code is a candidate chunk start (sliding sums
//                              from a block-wide prefix sum of the per-byte weights);
//   2. xhuff_parse_chunks_kernel (xhuff_decode.cuh, speculative mode) one lane per candidate parses ONE chunk into tokens and
//                              reports where it ended (next table) and how far its matches reach back before its own start;
//   3. the host stitches the chain from offset 0 (every link must land on a parsed candidate; anything else -- an incomplete
//      code, an error, a capacity problem -- sends the whole call to the serial decoder, which owns the status codes);
//   4. the token lists are executed: all chunks at once when no match reaches across a chunk start (streams made by this
//      library: its encoder resets the window per chunk), else in stream order by one block (xhs_apply_chain_kernel).
// The parse -- the 1,300-cycles-per-symbol part of a lone lane -- runs on as many lanes as the stream has chunks.
#pragma once
#include "common.cuh"
#include "lz_apply.cuh"

namespace msc {

constexpr int XHS_THREADS = 256;
constexpr int XHS_POS_PER_BLOCK = 1024;
constexpr int XHS_SPAN = XHS_POS_PER_BLOCK + 256;             // bytes a block looks at
constexpr int XHS_PER_THREAD = XHS_SPAN / XHS_THREADS;         // 5

__device__ __forceinline__ uint32_t xhs_weight(uint32_t b)
{
    const uint32_t lo = b & 15u, hi = b >> 4;
    return (lo ? 1u << (15u - lo) : 0u) + (hi ? 1u << (15u - hi) : 0u);
}

// cand[] receives (unordered) every offset s with s + 260 <= n whose 256 bytes form a complete prefix code; *ncand counts them
// (also beyond cap, so that the caller can tell an overflow).
__global__ void __launch_bounds__(XHS_THREADS)
xhs_candidates_kernel(const uint8_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ cand, uint32_t* __restrict__ ncand, uint32_t cap)
{
    __shared__ uint32_t pre[XHS_SPAN + 1];
    __shared__ uint32_t wsum[XHS_THREADS / 32];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t base = blockIdx.x * XHS_POS_PER_BLOCK; base + 260 <= n; base += gridDim.x * XHS_POS_PER_BLOCK) {
        uint32_t w[XHS_PER_THREAD], sum = 0;
#pragma unroll
        for (int k = 0; k < XHS_PER_THREAD; ++k) {
            const uint32_t at = base + (uint32_t)(tid * XHS_PER_THREAD + k);
            w[k] = at < n ? xhs_weight(in[at]) : 0u;
            sum += w[k];
        }
        const uint32_t incl = (uint32_t)warp_incl_scan((int)sum);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int k = 0; k < warp; ++k) run += wsum[k];
#pragma unroll
        for (int k = 0; k < XHS_PER_THREAD; ++k) { pre[tid * XHS_PER_THREAD + k] = run; run += w[k]; }
        if (tid == XHS_THREADS - 1) pre[XHS_SPAN] = run;
        __syncthreads();
        for (int p = tid; p < XHS_POS_PER_BLOCK; p += XHS_THREADS) {
            const uint32_t s = base + (uint32_t)p;
            if (s + 260u <= n && pre[p + 256] - pre[p] == 32768u) {
                const uint32_t at = atomicAdd(ncand, 1u);
                if (at < cap) cand[at] = s;
            }
        }
        __syncthreads();
    }
}

// The chunks of one stream in order (matches reach back across chunk starts, so the bytes must be produced in stream order): chunk j =
// tokens tok + slot[j] (ntok[j] of them), first output byte at pos[j] of `out`; the chunks are contiguous in the output.  One block
// with the last 128 KiB of output in a shared-memory ring (lz_apply_ring_run): a lone block would otherwise wait an L2 round trip
// per copy step.  Dynamic shared memory: LZR_RING bytes.
__global__ void __launch_bounds__(32 * LZB_WARPS)
xhs_apply_chain_kernel(const uint32_t* __restrict__ tok, const uint64_t* __restrict__ slot, const uint32_t* __restrict__ ntok,
                       const uint32_t* __restrict__ pos, uint32_t nchain, uint8_t* __restrict__ out)
{
    MSC_DYN_SMEM(ring_raw);
    __shared__ LzBlockSmem sm;
    uint8_t* ring = reinterpret_cast<uint8_t*>(ring_raw);
    uint32_t p = nchain ? pos[0] : 0u;
    for (uint32_t j = 0; j < nchain; ++j) p = lz_apply_ring_run(sm, ring, tok + slot[j], ntok[j], out, p);
}

// The same when no chunk reaches back before its own first byte: every chunk on its own warp.
__global__ void __launch_bounds__(32 * LZA_WARPS_PER_BLOCK)
xhs_apply_chunks_kernel(const uint32_t* __restrict__ tok, const uint64_t* __restrict__ slot, const uint32_t* __restrict__ ntok,
                        const uint32_t* __restrict__ pos, uint32_t nchain, uint8_t* __restrict__ out)
{
    const uint32_t warps_total = gridDim.x * LZA_WARPS_PER_BLOCK;
    for (uint32_t j = blockIdx.x * LZA_WARPS_PER_BLOCK + (threadIdx.x >> 5); j < nchain; j += warps_total)
        lz_apply_run(tok + slot[j], ntok[j], out + pos[j], 0u);
}

} // namespace msc
