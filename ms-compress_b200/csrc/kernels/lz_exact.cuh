// lz_exact.cuh -- the reference's match finder, position for position, as two bulk kernels.
//
// XpressDictionary (include/mscomp/XpressDictionary.h) is a hash-chain dictionary: Fill (:104-118) threads every position p of
// the input (p + 3 <= n) onto the chain of its 15-bit hash (:58-61), always a whole chunk ahead of the parse, and Find (:145-183)
// walks at most MaxChain = 11 links (:31, level 3) back from p while the link stays within MaxOffset, takes the first candidate
// that is strictly longer than what it has, and stops at NiceLength = 48.  Because Fill runs ahead of the parse, the link of a
// position is a pure function of the input:
//
//     prev(p) = the nearest q < p with hash(q) == hash(p)                                    (lz_prev_kernel)
//
// and so is the answer of Find at every position the parse could visit:
//
//     find(p) = first strict maximum of the match length over prev(p), prev(prev(p)), ... (<= 11 links, offsets <= window,
//               candidates whose first two bytes differ are hash collisions and only use up a link)       (lz_find_kernel)
//
// The serial part that remains is the greedy parse p += max(1, len(p)) over these answers (lz_front.cuh, lzf_parse_tile).
// Two quirks of the reference are reproduced because they change bytes:
//   * GetMatchLength (:74-93) never reports a match that reaches the last byte of the input: len <= n - p - 1.
//   * xpress_compress (src/xpress_compress.cpp:266-271) fills one 8 KiB chunk per loop iteration; after a match that jumps more
//     than a chunk ahead of the filled part, Find runs on positions that are not in the dictionary yet and finds nothing until the
//     fills have caught up -- see lzx_forced_literals below.
#pragma once
#include "common.cuh"
#include "lz_front.cuh"

namespace msc {

constexpr uint32_t LZX_SEG = 65536;            // work unit of the two kernels (positions); == XPE_SEG
constexpr uint32_t LZX_PIECE = 32768;          // lz_prev_kernel, 16-bit table: positions per work item (+ window of warm-up < 65535)
constexpr uint32_t LZX_CHAIN = 11, LZX_NICE = 48;
constexpr uint32_t LZX_CAP = 4096;             // lz_find_kernel measures at most this many bytes; len == LZX_CAP means "at least": the
                                               // parse extends the (rare) visited ones (lzx_extend), nobody measures the others
constexpr int LZX_FIND_THREADS = 256;

// the reference's progressive hash after three updates (XpressDictionary.h:58-61 with HashBits = 15): the first byte's high bits fall off
__device__ __forceinline__ uint32_t lzx_hash(uint32_t key3) { return (((key3 & 0xFFu) << 10) ^ ((key3 >> 3) & 0x1FE0u) ^ ((key3 >> 16) & 0xFFu)) & 0x7FFFu; }

// One warp per work item: `piece` positions of a unit plus `window` positions of warm-up before them.  T[h] = 1 + (position -
// warm-up start) of the latest position with hash h among those already passed, 0 = none; TE must hold piece + window.
// 32 positions per round: the nearest earlier lane with the same hash (__match_any_sync) wins, else the table; the highest lane of
// every hash group updates the table.  dist[] (one entry per position, indexed like the units' segments: seg_first[u] * LZX_SEG + p)
// receives p - prev(p), or 0 when there is none within the window.
template <typename TE>
__global__ void __launch_bounds__(32)
lz_prev_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
               const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ work_counter, uint16_t* __restrict__ dist,
               uint32_t window, uint32_t piece)
{
    MSC_DYN_SMEM(smem_raw);
    TE* T = reinterpret_cast<TE*>(smem_raw);
    const int lane = (int)threadIdx.x;
    const uint32_t per_seg = LZX_SEG / piece;
    const uint32_t total = seg_first[count] * per_seg;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= total) break;
        const uint32_t sgl = w / per_seg;
        uint32_t lo = 0, hi = count;                                  // the unit u with seg_first[u] <= sgl < seg_first[u + 1]
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seg_first[mid] <= sgl) lo = mid; else hi = mid; }
        const uint32_t u = lo, n = in_len[u];
        const uint32_t s0 = (sgl - seg_first[u]) * LZX_SEG + (w % per_seg) * piece;
        if (s0 >= n) continue;
        const uint32_t s1 = min(s0 + piece, n);
        const uint32_t w0 = s0 > window ? s0 - window : 0u;
        const LzfSrc S(in + in_off[u], n);
        uint16_t* dst = dist + (uint64_t)seg_first[u] * LZX_SEG;
        for (uint32_t i = (uint32_t)lane; i < 32768u * sizeof(TE) / 16u; i += 32) reinterpret_cast<uint4*>(T)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        for (uint32_t p0 = w0; p0 < s1; p0 += 32) {
            const uint32_t p = p0 + (uint32_t)lane;
            const bool on = p + 3 <= n;
            const uint32_t h = on ? lzx_hash(lzf_ld4(S, p)) : (0x8000u | (uint32_t)lane);
            const unsigned peers = __match_any_sync(FULL, h);
            const unsigned below = peers & ((1u << lane) - 1u);
            uint32_t d = 0;
            if (on) {
                if (below) d = (uint32_t)(lane - (31 - __clz((int)below)));
                else { const uint32_t t = T[h]; if (t) d = (p - w0 + 1u) - t; }
                if (d > window) d = 0;
            }
            __syncwarp();
            if (on && (peers >> lane) == 1u) T[h] = (TE)(p - w0 + 1u);
            __syncwarp();
            if (p >= s0 && p < s1) dst[p] = (uint16_t)d;
        }
        __syncwarp();
    }
}

// find(p) for every position of the units, one thread per position (persistent blocks pull 64 KiB segments).
// lenoff[] (indexed like dist[]) receives (len - 1) << 16 | offset, or 0 when Find would return less than 3.
// chunk = 0: the match may run to the end of the unit (plain Xpress); otherwise it is also clipped at the end of the chunk it
// starts in and needs 3 bytes of chunk left (Xpress-Huffman: src/xpress_huff_compress.cpp:90-93).
__global__ void __launch_bounds__(LZX_FIND_THREADS)
lz_find_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
               const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ work_counter, const uint16_t* __restrict__ dist,
               uint32_t* __restrict__ lenoff, uint32_t window, uint32_t chunk)
{
    __shared__ uint32_t s_work;
    const int tid = (int)threadIdx.x;
    const uint32_t total = seg_first[count];
    for (;;) {
        if (tid == 0) s_work = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t w = s_work;
        __syncthreads();
        if (w >= total) break;
        uint32_t lo = 0, hi = count;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seg_first[mid] <= w) lo = mid; else hi = mid; }
        const uint32_t u = lo, n = in_len[u];
        const uint32_t s0 = (w - seg_first[u]) * LZX_SEG, s1 = min(s0 + LZX_SEG, n);
        const uint8_t* src = in + in_off[u];
        const LzfSrc S(src, n);
        const uint16_t* dd = dist + (uint64_t)seg_first[u] * LZX_SEG;
        uint32_t* out = lenoff + (uint64_t)seg_first[u] * LZX_SEG;
        for (uint32_t p = s0 + (uint32_t)tid; p < s1; p += LZX_FIND_THREADS) {
            uint32_t res = 0;
            if (p + 3 <= n) {
                const uint32_t lim = n - p - 1u;                                 // GetMatchLength never reaches the last byte
                const uint32_t rem = chunk ? chunk - (p % chunk) : 0xFFFFFFFFu;  // Find itself does not know about chunks ...
                const uint32_t maxlen = min(lim, LZX_CAP);
                if (maxlen >= 3u && rem >= 3u) {
                    const uint32_t key = lzf_ld4(S, p);
                    uint32_t len = 2, off = 0, x = p;
                    for (uint32_t c = 0; c < LZX_CHAIN; ++c) {
                        const uint32_t d = dd[x];
                        if (!d || p - (x - d) > window) break;
                        x -= d;
                        const uint32_t cw = lzf_ld4(S, x);
                        if ((cw ^ key) & 0xFFFFu) continue;                     // a hash collision: uses up a link, nothing else
                        // a longer match must also agree on the four bytes ending at index len (bytes 0 and 1 are known to agree)
                        if (len >= 3u && lzf_ld4(S, x + len - 3u) != lzf_ld4(S, p + len - 3u)) continue;
                        const uint32_t l = lzf_match_len(S, src, p, x, maxlen);
                        if (l > len) {
                            len = l; off = p - x;
                            if (len >= LZX_NICE || len == maxlen) break;        // (nothing is longer than maxlen: the walk could only waste links)
                        }
                    }
                    len = min(len, rem);                                        // ... the caller clips what it found (:93)
                    if (len >= 3u) res = ((len - 1u) << 16) | off;
                }
            }
            out[p] = res;
        }
    }
}

// Number of positions after a match (start q, length len) of plain Xpress on which the reference's Find comes back empty-handed
// because the dictionary has not been filled that far (src/xpress_compress.cpp:269 fills ONE 8 KiB chunk per loop iteration; the
// stale window entries of an unfilled position lie more than MaxOffset back, XpressDictionary.h:164).  When q is a position where
// a match was found, the filled part ends at the next 8 KiB boundary after q (or n - 2); iteration j after the match stands at
// q + len + j with the filled part 8 KiB (j + 1) further: forced while q + len + j >= g + 8192 (j + 1).
__device__ __forceinline__ uint32_t lzx_forced_literals(uint32_t q, uint32_t len, uint32_t n)
{
    const uint32_t g = min(((q >> 13) + 1u) << 13, n - 2u);
    const uint32_t e = q + len;
    if (e < g + 8192u) return 0;
    return (e - g - 8192u) / 8191u + 1u;
}

// True length of a match at p with the given offset that lz_find_kernel measured up to LZX_CAP bytes: all threads of the block
// compare on from there, at most up to `maxlen`.  Must be called by all threads of the block; red = one shared word.
__device__ __forceinline__ uint32_t lzx_extend(const uint8_t* __restrict__ src, uint32_t p, uint32_t off, uint32_t maxlen, uint32_t* red)
{
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    uint32_t base = LZX_CAP;
    for (;;) {
        if (tid == 0) *red = 0xFFFFFFFFu;
        __syncthreads();
        const uint32_t i = base + tid;
        if (i < maxlen && src[p + i] != src[p + i - off]) atomicMin(red, i);
        __syncthreads();
        const uint32_t r = *red;
        __syncthreads();
        if (r != 0xFFFFFFFFu) return r;
        base += nt;
        if (base >= maxlen) return maxlen;
    }
}

} // namespace msc
