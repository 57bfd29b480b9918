// lz_jump.cuh -- LZ77 token execution of ONE stream with no order at all: pointer jumping.
//
// The in-order executors (lz_apply.cuh) produce the bytes of a stream front to back because a match may copy what the token before
// it has just written.  When a single stream has the GPU to itself (the one-shot calls) that order is the whole run time.  Here
// every output byte gets one 32-bit word: 0x80000000 | value once the byte is known, else the output position it is a copy of.
//   expand : literals are known at once; byte k of a match at position p with offset d points at p + k - d;
//   round  : word[i] = word[word[i]] for every unknown i, in place -- whatever a racing neighbour has already written there is
//            itself a valid state (a value, or a position further down the same chain), so no ordering is needed inside a round;
//            the distance to the literal at the end of a chain at least halves per round: ceil(log2(n)) rounds always suffice;
//   store  : the low bytes go to the output.
// Work is O(n log depth) instead of O(n), but on all SMs: 8 MiB take well under a millisecond where the in-order block takes 45 ms.
// Replaces the reference's byte-serial copy loops (src/xpress_huff_decompress.cpp:110-131, src/xpress_decompress.cpp:160-228) for
// that case; the tokens are the same ones the other executors take (lz_exec.cuh).
#pragma once
#include "common.cuh"
#include "lz_exec.cuh"

namespace msc {

constexpr int LZJ_THREADS = 256;
constexpr uint32_t LZJ_KNOWN = 0x80000000u;

// Chunk j of the chain: tokens tok + slot[j] (ntok[j] of them), first output byte at pos[j].  One block per chunk.
__global__ void __launch_bounds__(LZJ_THREADS)
lzj_expand_kernel(const uint32_t* __restrict__ tok, const uint64_t* __restrict__ slot, const uint32_t* __restrict__ ntok,
                  const uint32_t* __restrict__ pos, uint32_t nchain, uint32_t* __restrict__ word)
{
    __shared__ uint32_t wsum[LZJ_THREADS / 32];
    __shared__ uint32_t carry;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t j = blockIdx.x; j < nchain; j += gridDim.x) {
        const uint32_t* tk = tok + slot[j];
        const uint32_t n = ntok[j];
        if (tid == 0) carry = pos[j];
        __syncthreads();
        for (uint32_t t0 = 0; t0 < n; t0 += LZJ_THREADS) {
            const uint32_t t = (t0 + tid < n) ? tk[t0 + tid] : 0u;
            const bool is_match = (t >> 31) != 0;
            const uint32_t len = (t0 + tid < n) ? (is_match ? ((t >> 16) & 0x7FFFu) + 1u : 1u) : 0u;
            const uint32_t incl = (uint32_t)warp_incl_scan((int)len);
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t base = carry;
            for (int w = 0; w < warp; ++w) base += wsum[w];
            const uint32_t p = base + incl - len;
            if (len) {
                if (!is_match) word[p] = LZJ_KNOWN | (t & 0xFFu);
                else { const uint32_t from = p - (t & 0xFFFFu); for (uint32_t k = 0; k < len; ++k) word[p + k] = from + k; }
            }
            __syncthreads();
            if (tid == LZJ_THREADS - 1) carry = base + incl;
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(LZJ_THREADS)
lzj_round_kernel(uint32_t* __restrict__ word, uint32_t n)
{
    for (uint32_t i = blockIdx.x * LZJ_THREADS + threadIdx.x; i < n; i += gridDim.x * LZJ_THREADS) {
        const uint32_t s = word[i];
        if (s & LZJ_KNOWN) continue;
        word[i] = word[s];                                            // a value, or a position further down the chain
    }
}

// out[i] = the byte of word[i]; *unknown counts the words that are still positions (0 after enough rounds).
__global__ void __launch_bounds__(LZJ_THREADS)
lzj_store_kernel(const uint32_t* __restrict__ word, uint8_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ unknown)
{
    uint32_t bad = 0;
    for (uint32_t i = blockIdx.x * LZJ_THREADS + threadIdx.x; i < n; i += gridDim.x * LZJ_THREADS) {
        const uint32_t s = word[i];
        if (!(s & LZJ_KNOWN)) ++bad;
        out[i] = (uint8_t)s;
    }
    if (bad) atomicAdd(unknown, bad);
}

} // namespace msc
