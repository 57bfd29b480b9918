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
    __device__ __forceinline__ void new_chunk() {}
    __device__ __forceinline__ void before_symbol(uint32_t, uint32_t, uint32_t) {}
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

// ---- one BLOCK per stream ----------------------------------------------------------------------------------------------
// With few streams (plain Xpress: 1,024 buffers of 1 MiB) a warp per stream leaves the GPU idle and every round waits for the
// round before it.  Here the rounds of 32 tokens of ONE stream are pipelined over the LZB_WARPS warps of a block: warp w takes rounds
// w, w+LZB_WARPS, ... of a window of up to 512 rounds.  Round start positions come from a parallel prefix sum of the rounds' byte
// counts, literals are stored at once, and a match runs as soon as its source is final:
//   * the source lies in rounds that are completely done (below the frontier p0[done_upto], which advances in order), or
//   * the source lies in the match's own round and the watermark rule of that round allows it, or
//   * every earlier round is done, and the watermark rule of the round allows it (then this is lz_apply_run).
// The lowest unfinished round never waits for another round, so the pipeline cannot deadlock.
constexpr int LZB_WARPS = 8;
constexpr int LZB_ROUNDS = 512;
constexpr int LZB_EPT = LZB_ROUNDS / (32 * LZB_WARPS);   // rounds per thread in the prefix sum
static_assert(LZB_EPT * 32 * LZB_WARPS == LZB_ROUNDS, "window size");

struct LzBlockSmem {
    uint32_t p0[LZB_ROUNDS + 1];      // output position of the first byte of every round of the window (+ end)
    uint32_t flag[LZB_ROUNDS];        // round finished
    uint32_t done_upto;               // rounds [0, done_upto) of the window are finished
    uint32_t wsum[LZB_WARPS];
};

__device__ __forceinline__ uint32_t lz_apply_block_run(LzBlockSmem& sm, const uint32_t* __restrict__ tk, uint32_t n, uint8_t* __restrict__ o, uint32_t pstart)
{
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t w0 = 0; w0 < n; w0 += 32u * LZB_ROUNDS) {
        const uint32_t nt = min(n - w0, 32u * LZB_ROUNDS), R = (nt + 31) / 32;
        // bytes per round
        for (uint32_t r = warp; r < R; r += LZB_WARPS) {
            const uint32_t t = 32 * r + lane;
            const uint32_t tok = t < nt ? tk[w0 + t] : 0;
            const uint32_t len = t < nt ? ((tok >> 31) ? ((tok >> 16) & 0x7FFFu) + 1 : 1u) : 0u;
            const uint32_t sum = __reduce_add_sync(FULL, len);
            if (lane == 0) sm.p0[r] = sum;
        }
        __syncthreads();
        {   // exclusive scan of the (at most LZB_ROUNDS) round totals, LZB_EPT consecutive ones per thread
            uint32_t v[LZB_EPT], sum = 0;
#pragma unroll
            for (int i = 0; i < LZB_EPT; ++i) { const uint32_t r = (uint32_t)(LZB_EPT * tid + i); v[i] = r < R ? sm.p0[r] : 0u; sum += v[i]; }
            const uint32_t incl = (uint32_t)warp_incl_scan((int)sum);
            if (lane == 31) sm.wsum[warp] = incl;
            __syncthreads();
            uint32_t run = pstart + incl - sum;
            for (int w = 0; w < warp; ++w) run += sm.wsum[w];
#pragma unroll
            for (int i = 0; i < LZB_EPT; ++i) {
                const uint32_t r = (uint32_t)(LZB_EPT * tid + i);
                if (r < R) { sm.p0[r] = run; run += v[i]; if (r + 1 == R) sm.p0[R] = run; }
            }
            for (uint32_t r = tid; r < R; r += 32 * LZB_WARPS) sm.flag[r] = 0;
            if (tid == 0) sm.done_upto = 0;
        }
        __syncthreads();
        for (uint32_t r = warp; r < R; r += LZB_WARPS) {
            const uint32_t t = 32 * r + lane;
            const uint32_t tok = t < nt ? tk[w0 + t] : 0;
            const bool is_match = (tok >> 31) != 0;
            const uint32_t len = t < nt ? (is_match ? ((tok >> 16) & 0x7FFFu) + 1 : 1u) : 0u;
            const uint32_t off = tok & 0xFFFFu;
            const uint32_t p0r = sm.p0[r];
            const uint32_t pos = p0r + (uint32_t)warp_incl_scan((int)len) - len;
            if (len && !is_match) o[pos] = (uint8_t)tok;
            bool pending = is_match;
            const uint32_t a = pos - off, be = min(a + len, pos);        // the bytes a match reads: [a, be)
            for (;;) {
                __syncwarp();
                const unsigned W = __reduce_min_sync(FULL, pending ? pos : 0xffffffffu);
                if (W == 0xffffffffu) break;
                uint32_t d = 0;
                if (lane == 0) d = ld_shared_signal(&sm.done_upto);
                d = __shfl_sync(FULL, d, 0);
                const bool prev_done = d >= r;
                const uint32_t F = sm.p0[prev_done ? r : d];               // everything below F is final
                const bool can = pending && (be <= F || ((a >= p0r || prev_done) && (pos == W || be <= W)));
                if (!__any_sync(FULL, can)) { spin_pause(); continue; }    // all pending matches wait for earlier rounds
                if (!prev_done) __threadfence_block();                     // bytes of the rounds counted in d are visible from here on
                const uint32_t cl = can ? len : 0u;
                const uint32_t cincl = (uint32_t)warp_incl_scan((int)cl);
                const uint32_t total = __shfl_sync(FULL, cincl, 31);
                for (uint32_t b0 = 0; b0 < total; b0 += 32) {
                    const uint32_t bidx = b0 + lane;
                    int tt = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const uint32_t c = __shfl_sync(FULL, cincl, tt + step - 1);
                        if (c <= bidx) tt += step;
                    }
                    const uint32_t t_incl = __shfl_sync(FULL, cincl, tt), t_len = __shfl_sync(FULL, cl, tt);
                    const uint32_t t_pos = __shfl_sync(FULL, pos, tt), t_off = __shfl_sync(FULL, off, tt);
                    if (bidx < total) {
                        const uint32_t k = bidx - (t_incl - t_len);
                        const uint32_t sk = (t_off < t_len) ? (k % t_off) : k;
                        o[t_pos + k] = (uint8_t)ld_byte_cg(o + t_pos - t_off + sk);
                    }
                }
                if (can) pending = false;
            }
            // publish: the round is finished; whoever sees the frontier stuck at a finished round moves it
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {
                st_shared_signal(&sm.flag[r], 1u);
                __threadfence_block();
                for (;;) {
                    const uint32_t d = ld_shared_signal(&sm.done_upto);
                    if (d >= R || !ld_shared_signal(&sm.flag[d])) break;
                    atomicCAS(&sm.done_upto, d, d + 1);
                }
            }
        }
        __syncthreads();
        pstart = sm.p0[R];
        __syncthreads();
    }
    return pstart;
}

// The same pipeline for ONE stream that has the GPU to itself (the one-shot calls): the most recent LZR_RING bytes of the output
// also live in a shared-memory ring (byte x at ring[x & (LZR_RING - 1)]), and every match source is read from there -- a
// shared-memory round trip per copy step instead of one through L2, which is what a lone block waits for.  Offsets reach back at
// most 65,535 bytes, so the rounds in flight may span at most LZR_SPAN bytes: a window is cut where its rounds exceed that, and a
// single round that is longer by itself (very long matches) runs its tokens strictly in order, which keeps the writes monotone.
// Successive calls continue the same stream (pstart = where the previous call ended); the ring is carried over.
constexpr uint32_t LZR_RING = 128u << 10;
constexpr uint32_t LZR_SPAN = LZR_RING - 65536u - 4096u;
__device__ __forceinline__ uint32_t lz_apply_ring_run(LzBlockSmem& sm, uint8_t* __restrict__ ring, const uint32_t* __restrict__ tk, uint32_t n,
                                                      uint8_t* __restrict__ o, uint32_t pstart)
{
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t RM = LZR_RING - 1u;
    for (uint32_t w0 = 0; w0 < n; ) {
        const uint32_t nt_all = min(n - w0, 32u * LZB_ROUNDS);
        uint32_t nt = nt_all, R = (nt + 31) / 32;
        // bytes per round
        for (uint32_t r = warp; r < R; r += LZB_WARPS) {
            const uint32_t t = 32 * r + lane;
            const uint32_t tok = t < nt ? tk[w0 + t] : 0;
            const uint32_t len = t < nt ? ((tok >> 31) ? ((tok >> 16) & 0x7FFFu) + 1 : 1u) : 0u;
            const uint32_t sum = __reduce_add_sync(FULL, len);
            if (lane == 0) sm.p0[r] = sum;
        }
        __syncthreads();
        {   // exclusive scan of the (at most LZB_ROUNDS) round totals, LZB_EPT consecutive ones per thread
            uint32_t v[LZB_EPT], sum = 0;
#pragma unroll
            for (int i = 0; i < LZB_EPT; ++i) { const uint32_t r = (uint32_t)(LZB_EPT * tid + i); v[i] = r < R ? sm.p0[r] : 0u; sum += v[i]; }
            const uint32_t incl = (uint32_t)warp_incl_scan((int)sum);
            if (lane == 31) sm.wsum[warp] = incl;
            __syncthreads();
            uint32_t run = pstart + incl - sum;
            for (int w = 0; w < warp; ++w) run += sm.wsum[w];
#pragma unroll
            for (int i = 0; i < LZB_EPT; ++i) {
                const uint32_t r = (uint32_t)(LZB_EPT * tid + i);
                if (r < R) { sm.p0[r] = run; run += v[i]; if (r + 1 == R) sm.p0[R] = run; }
            }
            for (uint32_t r = tid; r < R; r += 32 * LZB_WARPS) sm.flag[r] = 0;
            if (tid == 0) sm.done_upto = 0;
        }
        __syncthreads();
        // cut the window where the rounds in flight would span more than the ring can hold behind a 64 KiB offset
        bool serial = false;
        {
            int fit = 0;
            for (uint32_t r = tid; r < R; r += 32 * LZB_WARPS) fit += (sm.p0[r + 1] - pstart <= LZR_SPAN) ? 1 : 0;
            fit = __reduce_add_sync(FULL, fit);
            if (lane == 0) sm.wsum[warp] = (uint32_t)fit;
            __syncthreads();
            uint32_t total_fit = 0;
            for (int w = 0; w < LZB_WARPS; ++w) total_fit += sm.wsum[w];
            __syncthreads();
            if (total_fit == 0) { serial = true; total_fit = 1; }
            if (total_fit < R) { R = total_fit; nt = 32u * R; }
        }
        for (uint32_t r = warp; r < R; r += LZB_WARPS) {
            const uint32_t t = 32 * r + lane;
            const uint32_t tok = t < nt ? tk[w0 + t] : 0;
            const bool is_match = (tok >> 31) != 0;
            const uint32_t len = t < nt ? (is_match ? ((tok >> 16) & 0x7FFFu) + 1 : 1u) : 0u;
            const uint32_t off = tok & 0xFFFFu;
            const uint32_t p0r = sm.p0[r];
            const uint32_t pos = p0r + (uint32_t)warp_incl_scan((int)len) - len;
            if (len && !is_match) { o[pos] = (uint8_t)tok; ring[pos & RM] = (uint8_t)tok; }
            bool pending = is_match;
            const uint32_t a = pos - off, be = min(a + len, pos);        // the bytes a match reads: [a, be)
            for (;;) {
                __syncwarp();
                const unsigned W = __reduce_min_sync(FULL, pending ? pos : 0xffffffffu);
                if (W == 0xffffffffu) break;
                uint32_t d = 0;
                if (lane == 0) d = ld_shared_signal(&sm.done_upto);
                d = __shfl_sync(FULL, d, 0);
                const bool prev_done = d >= r;
                const uint32_t F = sm.p0[prev_done ? r : d];               // everything below F is final
                const bool can = pending && (serial ? (prev_done && pos == W) : (be <= F || ((a >= p0r || prev_done) && (pos == W || be <= W))));
                if (!__any_sync(FULL, can)) { spin_pause(); continue; }    // all pending matches wait for earlier rounds
                if (!prev_done) __threadfence_block();                     // bytes of the rounds counted in d are visible from here on
                const uint32_t cl = can ? len : 0u;
                const uint32_t cincl = (uint32_t)warp_incl_scan((int)cl);
                const uint32_t total = __shfl_sync(FULL, cincl, 31);
                for (uint32_t b0 = 0; b0 < total; b0 += 32) {
                    const uint32_t bidx = b0 + lane;
                    int tt = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const uint32_t c = __shfl_sync(FULL, cincl, tt + step - 1);
                        if (c <= bidx) tt += step;
                    }
                    const uint32_t t_incl = __shfl_sync(FULL, cincl, tt), t_len = __shfl_sync(FULL, cl, tt);
                    const uint32_t t_pos = __shfl_sync(FULL, pos, tt), t_off = __shfl_sync(FULL, off, tt);
                    if (bidx < total) {
                        const uint32_t k = bidx - (t_incl - t_len);
                        const uint32_t sk = (t_off < t_len) ? (k % t_off) : k;
                        const uint8_t v = ring[(t_pos - t_off + sk) & RM];
                        ring[(t_pos + k) & RM] = v;
                        o[t_pos + k] = v;
                    }
                }
                if (can) pending = false;
            }
            // publish: the round is finished; whoever sees the frontier stuck at a finished round moves it
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {
                st_shared_signal(&sm.flag[r], 1u);
                __threadfence_block();
                for (;;) {
                    const uint32_t d = ld_shared_signal(&sm.done_upto);
                    if (d >= R || !ld_shared_signal(&sm.flag[d])) break;
                    atomicCAS(&sm.done_upto, d, d + 1);
                }
            }
        }
        __syncthreads();
        pstart = sm.p0[R];
        __syncthreads();
        w0 += nt;
    }
    return pstart;
}

// Stream i: tokens tok_base + 4 * out_off[i] (ntok[i] of them) -> out + out_off[i], one block per stream.
__global__ void __launch_bounds__(32 * LZB_WARPS)
lz_apply_block_kernel(const uint32_t* __restrict__ tok_base, const uint32_t* __restrict__ ntok, const uint64_t* __restrict__ out_off,
                      const int32_t* __restrict__ status, uint32_t count, uint8_t* __restrict__ out)
{
    __shared__ LzBlockSmem sm;
    for (uint32_t s = blockIdx.x; s < count; s += gridDim.x) {
        if (status[s] < 0) continue;                                   // (block-uniform)
        lz_apply_block_run(sm, tok_base + out_off[s], ntok[s], out + out_off[s], 0u);
    }
}

} // namespace msc
