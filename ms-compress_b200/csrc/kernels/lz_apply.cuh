// lz_apply.cuh -- second stage of the two-kernel Xpress / Xpress-Huffman decoders: LZ77 token execution,
// one WARP per stream.
//
// Stage one (xhuff_decode.cuh / xpress_decode.cuh with the LaneTok sink) is the serial part of the formats -- bit and
// flag parsing -- done by one lane per stream and free of random memory traffic: it only appends 4-byte tokens to a
// per-stream list.  Stage two is this kernel: the tokens of a stream are known, so a warp takes 32 of them at a
// time: a prefix sum of their lengths gives every token its output position, literals are stored in parallel, and
// the matches are copied under the watermark rule of the LZNT1 decoder (a pending match may run once its whole
// source lies below the lowest pending match destination of the warp, or when it is that lowest one).  All loads of
// a round are in flight together and the stores of a stream are contiguous, which is what the one-lane-per-stream
// output path cannot offer.  Replaces the copy loops / FAST_COPY of the reference (include/mscomp/internal.h:568-598,
// src/xpress_huff_decompress.cpp:112-125, src/xpress_decompress.cpp:440-453).
#pragma once
#include "common.cuh"
#include "lz_exec.cuh"
#include "lane_out.cuh"

namespace msc {

// token: bit31 = match; literal: low 8 bits; match: offset in [0,16), (len-1) in [16,31)
__device__ __forceinline__ uint32_t lzt_lit(uint32_t b) { return b & 0xFFu; }
__device__ __forceinline__ uint32_t lzt_match(uint32_t off, uint32_t len) { return 0x80000000u | ((len - 1) << 16) | off; }

// Per-lane token appender (stage one sink).  Mirrors the interface of LaneOut so the parsers are written once.
struct LaneTok {
    uint32_t* tok;          // this stream's token list
    uint32_t ntok;          // tokens written
    uint32_t opos;          // output bytes the tokens stand for
    __device__ __forceinline__ void init(uint32_t* t) { tok = t; ntok = 0; opos = 0; }
    __device__ __forceinline__ void put_byte(uint32_t v) { tok[ntok++] = lzt_lit(v); ++opos; }
    __device__ __forceinline__ void copy_match(uint32_t off, uint32_t len)
    {
        opos += len;
        while (len) {                                  // (offsets are <= 65535 in both formats)
            const uint32_t l = len > LZQ_MAXLEN ? LZQ_MAXLEN : len;
            tok[ntok++] = lzt_match(off, l);
            len -= l;
        }
    }
    // one item, branch-free in the common case (a match of at most LZQ_MAXLEN bytes, or a literal)
    __device__ __forceinline__ void emit(bool is_match, uint32_t lit, uint32_t off, uint32_t len)
    {
        if (is_match && len > LZQ_MAXLEN) { copy_match(off, len); return; }
        tok[ntok++] = is_match ? lzt_match(off, len) : lzt_lit(lit);
        opos += len;
    }
    __device__ __forceinline__ void finish() {}
};

// sink plumbing shared by the two parsers
__device__ __forceinline__ void sink_init(LaneOut& O, uint8_t* out, uint32_t*, uint64_t off) { O.init(out + off); }
__device__ __forceinline__ void sink_init(LaneTok& O, uint8_t*, uint32_t* tok, uint64_t off) { O.init(tok + off); }
__device__ __forceinline__ void sink_count(const LaneOut&, uint32_t*, uint32_t) {}
__device__ __forceinline__ void sink_count(const LaneTok& O, uint32_t* ntok, uint32_t i) { ntok[i] = O.ntok; }

constexpr int LZA_WARPS_PER_BLOCK = 4;

// Executes n tokens (tk[0..n)) of one stream whose first token starts at output position p0 of o; returns the position
// after the last one.  Whole warp.
__device__ __forceinline__ uint32_t lz_apply_run(const uint32_t* __restrict__ tk, uint32_t n, uint8_t* __restrict__ o, uint32_t p0)
{
    const int lane = lane_id();
    for (uint32_t t0 = 0; t0 < n; t0 += 32) {
        const uint32_t tok = (t0 + lane < n) ? tk[t0 + lane] : 0;
        const bool is_match = (tok >> 31) != 0;
        const uint32_t len = (t0 + lane < n) ? (is_match ? ((tok >> 16) & 0x7FFFu) + 1 : 1u) : 0u;
        const uint32_t off = tok & 0xFFFFu;
        const uint32_t incl = (uint32_t)warp_incl_scan((int)len);
        const uint32_t pos = p0 + incl - len;
        if (len && !is_match) o[pos] = (uint8_t)tok;
        bool pending = is_match;
        for (;;) {
            __syncwarp();
            const unsigned W = __reduce_min_sync(FULL, pending ? pos : 0xffffffffu);
            if (W == 0xffffffffu) break;
            const bool can = pending && (pos == W || pos - off + len <= W);
            // The runnable matches are copied BYTE-transposed: their bytes are numbered consecutively, lane j takes
            // byte j, j+32, ... -- neighbouring lanes write neighbouring addresses and read neighbouring sources.
            const uint32_t cl = can ? len : 0u;
            const uint32_t cincl = (uint32_t)warp_incl_scan((int)cl);
            const uint32_t total = __shfl_sync(FULL, cincl, 31);
            for (uint32_t b0 = 0; b0 < total; b0 += 32) {
                const uint32_t bidx = b0 + lane;
                // owner token: first lane whose inclusive count exceeds bidx (binary search over the lanes)
                int t = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const uint32_t c = __shfl_sync(FULL, cincl, t + step - 1);
                    if (c <= bidx) t += step;
                }
                const uint32_t t_incl = __shfl_sync(FULL, cincl, t), t_len = __shfl_sync(FULL, cl, t);
                const uint32_t t_pos = __shfl_sync(FULL, pos, t), t_off = __shfl_sync(FULL, off, t);
                if (bidx < total) {
                    const uint32_t k = bidx - (t_incl - t_len);                      // byte index inside the match
                    // overlapped match (off < len): byte k equals byte k mod off of the period just before it
                    const uint32_t sk = (t_off < t_len) ? (k % t_off) : k;
                    o[t_pos + k] = o[t_pos - t_off + sk];
                }
            }
            if (can) pending = false;
        }
        p0 += __shfl_sync(FULL, incl, 31);
    }
    return p0;
}

// Stream i: tokens tok_base + 4 * out_off[i] (ntok[i] of them) -> out + out_off[i]; skipped when status[i] < 0.
__global__ void __launch_bounds__(32 * LZA_WARPS_PER_BLOCK)
lz_apply_kernel(const uint32_t* __restrict__ tok_base, const uint32_t* __restrict__ ntok, const uint64_t* __restrict__ out_off,
                const int32_t* __restrict__ status, uint32_t count, uint8_t* __restrict__ out)
{
    const uint32_t warps_total = gridDim.x * LZA_WARPS_PER_BLOCK;
    for (uint32_t s = blockIdx.x * LZA_WARPS_PER_BLOCK + (threadIdx.x >> 5); s < count; s += warps_total) {
        if (status[s] < 0) continue;
        lz_apply_run(tok_base + out_off[s], ntok[s], out + out_off[s], 0u);     // 4 bytes of token space per output byte
    }
}

} // namespace msc
