// lz_front.cuh -- LZ77 front end shared by the Xpress and Xpress-Huffman encoders (one thread block per unit).
//
// Stands where the reference has XpressDictionary::Fill/Find (include/mscomp/XpressDictionary.h:104-118, :145-183)
// and the greedy loops of xh_compress_lz77 (src/xpress_huff_compress.cpp:52-153) / xpress_compress
// (src/xpress_compress.cpp:266-325).  It is NOT a restatement of that matcher: the reference inserts positions
// one by one into hash chains (a serial structure); here match finding is bulk-synchronous:
//
//   for every step of 256 consecutive positions (one per thread):
//     lookup  : candidate A = most recent earlier position with the same 3-byte hash seen in PREVIOUS steps
//               (shared-memory table), candidate B = nearest earlier lane of the same warp with the same hash
//               (__match_any_sync); both are verified / extended against the input itself
//     insert  : atomicMax of (position+1) into the table          (deterministic: max = most recent)
//   every 2048 positions: the greedy parse p += max(1, len[p]) -- successor links nxt(q) for all match positions in
//   parallel, then one thread follows the chain -- leaves a "visited" bitmap; ranks inside it give token indices.
//
// Output bytes differ from the reference's (different candidate set); any greedy LZ77 parse with valid
// (offset, length) pairs is a valid stream for the reference's decoders.
#pragma once
#include "common.cuh"
#include "lz_exec.cuh"   // token format, LZQ_MAXLEN

namespace msc {

constexpr int LZF_THREADS = 256;
constexpr int LZF_TILE = 2048;
constexpr int LZF_HBITS = 13;
constexpr uint32_t LZF_TOK_EOS = 0x40000000u;     // literal-class token standing for the end-of-stream symbol

struct LzFrontSmem {
    uint32_t H[1 << LZF_HBITS];        // position+1 of the most recent earlier position per hash (0 = none)
    uint32_t lenoff[LZF_TILE];         // (len-1) << 16 | offset; 0 = no match at this position
    uint32_t M[LZF_TILE / 32];         // match bitmap of the tile
    uint32_t V[LZF_TILE / 32];         // visited (= token start) bitmap of the tile
    uint32_t prefV[LZF_TILE / 32];     // exclusive prefix popcount of V
    uint32_t K[LZF_TILE / 32];         // visited matches (marks of the pointer-jumping parse)
    uint16_t nxt[LZF_TILE];            // nxt(q) = first match position >= q + len[q] (relative; LZF_TILE = leaves the tile)
    uint16_t nxt2[LZF_TILE], nxt4[LZF_TILE];   // nxt o nxt and its square
    uint16_t chain[LZF_TILE / 12 + 8]; // every 4th visited match of the tile
    uint32_t nchain;
    uint32_t next_out;                 // furthest match end seen in this tile (relative)
    uint32_t next_p;                   // position (relative to the unit) where parsing resumes
    uint32_t ntok;                     // tokens emitted so far in the unit
};

// 4 bytes at src[p] (little endian), bytes at or beyond n read as 0; src need not be aligned.
__device__ __forceinline__ uint32_t lzf_ld4(const uint8_t* __restrict__ src, uint32_t p, uint32_t n)
{
    if (p + 4 <= n) {
        const uint8_t* a = src + p;
        const uint32_t sh = (uint32_t)((uintptr_t)a & 3u);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(a - sh);
        if (sh == 0) return __ldg(w);
        // the second word holds at least one byte of [p, p+4): it exists
        return __funnelshift_r(__ldg(w), __ldg(w + 1), 8 * sh);
    }
    uint32_t v = 0;
    for (uint32_t i = 0; i < 4 && p + i < n; ++i) v |= (uint32_t)src[p + i] << (8 * i);
    return v;
}
__device__ __forceinline__ uint32_t lzf_hash3(uint32_t x) { return ((x & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - LZF_HBITS); }

// common prefix length of src[p..] and src[c..] (c < p), at most maxlen
__device__ __forceinline__ uint32_t lzf_match_len(const uint8_t* __restrict__ src, uint32_t p, uint32_t c, uint32_t maxlen, uint32_t n)
{
    uint32_t i = 0;
    while (i + 4 <= maxlen) {
        const uint32_t d = lzf_ld4(src, p + i, n) ^ lzf_ld4(src, c + i, n);
        if (d) return i + ((uint32_t)(__ffs((int)d) - 1) >> 3);
        i += 4;
    }
    while (i < maxlen && src[p + i] == src[c + i]) ++i;
    return i;
}

// Match search + greedy parse for tile [t0, t0 + LZF_TILE) of the unit src[0..n).  `window` = largest usable
// offset, `far3` = largest offset accepted for a 3-byte match.  On return sm.lenoff / sm.M / sm.V / sm.prefV
// describe the tile and sm.next_p has advanced; the number of tokens starting in the tile is returned.
// Must be called by all LZF_THREADS threads.
__device__ __forceinline__ uint32_t lzf_tile(LzFrontSmem& sm, const uint8_t* __restrict__ src, uint32_t n, uint32_t t0,
                                             uint32_t window, uint32_t far3)
{
    const int tid = (int)threadIdx.x, lane = tid & 31;
    for (uint32_t s0 = t0; s0 < t0 + LZF_TILE; s0 += LZF_THREADS) {
        const uint32_t p = s0 + tid;
        uint32_t res = 0, h = 0xFFFFFFFFu;
        const bool on = p + 3 <= n;
        uint32_t key = 0;
        if (on) { key = lzf_ld4(src, p, n) & 0xFFFFFFu; h = lzf_hash3(key); }
        const unsigned peers = __match_any_sync(FULL, on ? h : (0x80000000u | (uint32_t)lane));
        if (on) {
            const uint32_t maxlen = min(n - p, LZQ_MAXLEN);
            uint32_t best = 0, boff = 0;
            const unsigned below = peers & ((1u << lane) - 1);
            if (below) {                                            // candidate B: nearest earlier lane, same hash
                const uint32_t d = (uint32_t)(lane - (31 - __clz((int)below)));
                if (d <= window && (lzf_ld4(src, p - d, n) & 0xFFFFFFu) == key) { best = lzf_match_len(src, p, p - d, maxlen, n); boff = d; }
            }
            const uint32_t a = sm.H[h];
            if (a) {                                                // candidate A: most recent position of earlier steps
                const uint32_t c = a - 1, d = p - c;
                if (d <= window && d != boff && best < maxlen && (lzf_ld4(src, c, n) & 0xFFFFFFu) == key) {
                    const uint32_t l = lzf_match_len(src, p, c, maxlen, n);
                    if (l > best) { best = l; boff = d; }
                }
            }
            if (best >= 3 && !(best == 3 && boff > far3)) res = ((best - 1) << 16) | boff;
        }
        sm.lenoff[p - t0] = res;
        const unsigned mm = __ballot_sync(FULL, res != 0);
        if (lane == 0) sm.M[(p - t0) >> 5] = mm;
        __syncthreads();
        if (on) atomicMax(&sm.H[h], p + 1);
        __syncthreads();
    }
    // Greedy parse of the tile, p += max(1, len[p]): for a match position q the parse continues at
    // nxt(q) = first match position >= q + len[q] (computed for all match positions in parallel); the visited matches
    // are the chain from the first match at or after the carried-in position -- one shared-memory load per link.
    {
        const uint32_t tend_r = min(t0 + LZF_TILE, n) - t0;                 // tile length (relative)
        const uint32_t start_r = sm.next_p > t0 ? sm.next_p - t0 : 0;       // >= tend_r when a long match skips the tile
        auto first_match_from = [&](uint32_t e) -> uint32_t {               // relative position or LZF_TILE
            if (e >= (uint32_t)LZF_TILE) return LZF_TILE;
            uint32_t w = e >> 5;
            uint32_t bits = sm.M[w] & (0xffffffffu << (e & 31));
            while (!bits && ++w < LZF_TILE / 32) bits = sm.M[w];
            return bits ? (w * 32 + (uint32_t)__ffs((int)bits) - 1) : (uint32_t)LZF_TILE;
        };
        auto or_range = [&](uint32_t lo, uint32_t hi) {                     // V |= [lo, hi)
            for (uint32_t w = lo >> 5; lo < hi && w <= ((hi - 1) >> 5); ++w) {
                const uint32_t lo_b = max(lo, w * 32) - w * 32, hi_b = min(hi, w * 32 + 32) - w * 32;
                atomicOr(&sm.V[w], (hi_b == 32 ? 0xffffffffu : ((1u << hi_b) - 1)) & (0xffffffffu << lo_b));
            }
        };
        uint16_t* cur = sm.nxt;
        for (int r = tid; r < LZF_TILE; r += LZF_THREADS)
            cur[r] = (uint16_t)(((sm.M[r >> 5] >> (r & 31)) & 1) ? first_match_from((uint32_t)r + (sm.lenoff[r] >> 16) + 1) : (uint32_t)LZF_TILE);
        if (tid < LZF_TILE / 32) { sm.V[tid] = 0; sm.K[tid] = 0; }
        if (tid == 0) sm.next_out = 0;
        __syncthreads();
        if (tid == 32 && start_r < tend_r) or_range(start_r, min(first_match_from(start_r), tend_r));   // leading literals
        // two rounds of pointer doubling (nxt2 = nxt o nxt, nxt4 = nxt2 o nxt2) cut the serial part of the chase by four
        for (int r = tid; r < LZF_TILE; r += LZF_THREADS) { const uint32_t t = cur[r]; sm.nxt2[r] = (uint16_t)(t < (uint32_t)LZF_TILE ? cur[t] : (uint32_t)LZF_TILE); }
        __syncthreads();
        for (int r = tid; r < LZF_TILE; r += LZF_THREADS) { const uint32_t t = sm.nxt2[r]; sm.nxt4[r] = (uint16_t)(t < (uint32_t)LZF_TILE ? sm.nxt2[t] : (uint32_t)LZF_TILE); }
        __syncthreads();
        if (tid == 0) {                             // every 4th visited match: one shared-memory load per link
            uint32_t k = 0;
            if (start_r < tend_r) {
                uint32_t q = first_match_from(start_r);
                while (q < tend_r) { sm.chain[k++] = (uint16_t)q; q = sm.nxt4[q]; }
            }
            sm.nchain = k;
        }
        __syncthreads();
        for (uint32_t k = tid; k < sm.nchain; k += LZF_THREADS) {                      // the three in between, in parallel
            uint32_t q = sm.chain[k];
            atomicOr(&sm.K[q >> 5], 1u << (q & 31));
            const uint32_t q2 = sm.nxt2[q];
            q = cur[q];
            if (q < tend_r) atomicOr(&sm.K[q >> 5], 1u << (q & 31));
            if (q2 < tend_r) {
                atomicOr(&sm.K[q2 >> 5], 1u << (q2 & 31));
                const uint32_t q3 = cur[q2];
                if (q3 < tend_r) atomicOr(&sm.K[q3 >> 5], 1u << (q3 & 31));
            }
        }
        __syncthreads();
        // visited = marked matches + the literals up to the next marked match; carry-out = furthest match end
        for (int r = tid; r < LZF_TILE; r += LZF_THREADS) {
            if ((sm.K[r >> 5] >> (r & 31)) & 1) {
                const uint32_t e = (uint32_t)r + (sm.lenoff[r] >> 16) + 1;
                or_range((uint32_t)r, (uint32_t)r + 1);
                or_range(e, min(first_match_from(e), tend_r));
                atomicMax(&sm.next_out, e);
            }
        }
        __syncthreads();
        if (tid == 0 && start_r < tend_r) sm.next_p = t0 + max(sm.next_out, tend_r);
    }
    __syncthreads();
    if (tid < 32) {     // exclusive prefix popcount of V (64 words, 2 per lane)
        const uint32_t a = (uint32_t)__popc(sm.V[2 * lane]), b = (uint32_t)__popc(sm.V[2 * lane + 1]);
        const uint32_t incl = (uint32_t)warp_incl_scan((int)(a + b));
        sm.prefV[2 * lane] = incl - a - b;
        sm.prefV[2 * lane + 1] = incl - b;
    }
    __syncthreads();
    // every thread needs the tile's token count: recompute from the last prefix entry
    return sm.prefV[LZF_TILE / 32 - 1] + (uint32_t)__popc(sm.V[LZF_TILE / 32 - 1]);
}

} // namespace msc
