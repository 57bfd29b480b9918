// lz_exec.cuh -- shared pieces of the "one lane per stream" decoders (Xpress-Huffman, plain Xpress).
//
// Both formats are bit/byte-serial per stream, so the parallel axis is the batch: every
// lane of a warp owns one independent buffer and parses it into a small token queue in shared memory; the warp
// then executes each lane's queue cooperatively (prefix sum of lengths -> output positions, literals stored in
// parallel, matches copied under the same watermark rule as the LZNT1 decoder, with the history read from the
// output buffer in global memory / L2).  This replaces the reference's byte-serial copy loops and FAST_COPY
// (include/mscomp/internal.h:568-598, src/xpress_huff_decompress.cpp:112-125, src/xpress_decompress.cpp:440-453).
#pragma once
#include "common.cuh"

namespace msc {

constexpr int LZQ_TOKENS = 32;                 // tokens per lane per round
constexpr int LZQ_STRIDE = LZQ_TOKENS + 1;     // +1: conflict-free for both the per-lane writes and the per-stream reads
constexpr uint32_t LZQ_MAXLEN = 32768;         // longest single match token; longer matches are split

// token: bit31 = match; literal: low 8 bits; match: offset in [0,16), (len-1) in [16,31)
__device__ __forceinline__ uint32_t lzq_lit(uint32_t b) { return b & 0xFFu; }
__device__ __forceinline__ uint32_t lzq_match(uint32_t off, uint32_t len) { return 0x80000000u | ((len - 1) << 16) | off; }

// Byte-sequential reader over global memory with a 16-byte window held in registers (two aligned 8-byte words).
// Reads of <= 4 bytes at the cursor never touch memory except when the cursor crosses an 8-byte boundary.
struct ByteWindow {
    const uint8_t* base;     // aligned-down address of w0
    uint64_t w0, w1;
    const uint8_t* lim;      // first 8-byte aligned address at or beyond the end of the buffer
    __device__ __forceinline__ void init(const uint8_t* p, const uint8_t* end)
    {
        base = p - ((uintptr_t)p & 7u);
        lim = end + ((8u - ((uintptr_t)end & 7u)) & 7u);
        w0 = (base < lim) ? *reinterpret_cast<const uint64_t*>(base) : 0;
        w1 = (base + 8 < lim) ? *reinterpret_cast<const uint64_t*>(base + 8) : 0;
    }
    // make [p, p+4) readable from the window; cheap when p moves forward by a few bytes
    __device__ __forceinline__ void seek(const uint8_t* p)
    {
        if (p < base) { init(p, lim); return; }                 // rewind (second pass over a code-length table)
        while (p >= base + 8) {
            if (p >= base + 16) { init(p, lim); return; }       // long jump (new chunk table etc.)
            base += 8; w0 = w1;
            w1 = (base + 8 < lim) ? *reinterpret_cast<const uint64_t*>(base + 8) : 0;
        }
    }
    __device__ __forceinline__ uint32_t get32(const uint8_t* p)
    {   // 4 bytes at p (little endian), p in [base, base+8)
        const uint32_t sh = (uint32_t)(p - base) * 8;            // 0..56
        uint64_t v = w0 >> sh;
        if (sh > 32) v |= w1 << (64 - sh);
        return (uint32_t)v;
    }
};

// Execute the queued tokens of every lane's stream.  q: [32][LZQ_STRIDE] tokens; ntok / out / opos are per-lane
// values of the lane's own stream; on return opos has advanced.  All validity checks were done by the parser.
__device__ __forceinline__ void lzq_execute(const uint32_t* __restrict__ q, int ntok, uint8_t* out, uint32_t& opos)
{
    const int lane = lane_id();
    const unsigned have = __ballot_sync(FULL, ntok > 0);
    for (unsigned rest = have; rest; rest &= rest - 1) {
        const int s = __ffs((int)rest) - 1;
        const int n = __shfl_sync(FULL, ntok, s);
        uint8_t* o = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, (unsigned long long)(uintptr_t)out, s));
        const uint32_t p0 = __shfl_sync(FULL, opos, s);
        const uint32_t tok = (lane < n) ? q[s * LZQ_STRIDE + lane] : 0;
        const bool is_match = (tok >> 31) != 0;
        const uint32_t len = (lane < n) ? (is_match ? ((tok >> 16) & 0x7FFFu) + 1 : 1u) : 0u;
        const uint32_t off = tok & 0xFFFFu;
        const uint32_t incl = (uint32_t)warp_incl_scan((int)len);
        const uint32_t pos = p0 + incl - len;
        if (lane < n && !is_match) o[pos] = (uint8_t)tok;
        bool pending = is_match;
        for (;;) {
            __syncwarp();
            const unsigned W = __reduce_min_sync(FULL, pending ? pos : 0xffffffffu);
            if (W == 0xffffffffu) break;
            const bool can = pending && (pos == W || pos - off + len <= W);
            unsigned coop = __ballot_sync(FULL, can && len >= 32);
            if (can && len < 32) {
                if (off >= len) {
                    // no overlap: batch the loads ahead of the stores (8 at a time) for memory-level parallelism
                    for (uint32_t i = 0; i < len; i += 8) {
                        uint8_t b[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) b[k] = (i + k < len) ? o[pos - off + i + k] : (uint8_t)0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) if (i + k < len) o[pos + i + k] = b[k];
                    }
                } else {
                    for (uint32_t i = 0; i < len; ++i) o[pos + i] = o[pos - off + i];
                }
            }
            while (coop) {
                const int src = __ffs((int)coop) - 1;
                coop &= coop - 1;
                const uint32_t d = __shfl_sync(FULL, pos, src), l = __shfl_sync(FULL, len, src), f = __shfl_sync(FULL, off, src);
                if (f >= 32) {
                    for (uint32_t i = 0; i < l; i += 32) {
                        if (i + lane < l) o[d + i + lane] = o[d - f + i + lane];
                        __syncwarp();
                    }
                } else {
                    for (uint32_t i = lane; i < l; i += 32) o[d + i] = o[d - f + (i % f)];
                    __syncwarp();
                }
            }
            if (can) pending = false;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (lane == s) opos = p0 + total;
    }
    __syncwarp();
}

} // namespace msc
