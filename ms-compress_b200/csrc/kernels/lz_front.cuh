// lz_front.cuh -- LZ77 front end shared by the Xpress and Xpress-Huffman encoders (one thread block per unit).
//
// Stands where the reference has XpressDictionary::Fill/Find (include/mscomp/XpressDictionary.h:104-118, :145-183)
// and the greedy loops of xh_compress_lz77 (src/xpress_huff_compress.cpp:52-153) / xpress_compress
// (src/xpress_compress.cpp:266-325).  It is NOT a restatement of that matcher: the reference inserts positions
// one by one into hash chains (a serial structure); here match finding is bulk-synchronous:
//
//   for every step of 512 consecutive positions (two per thread):
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
constexpr int LZF_STEP = 2;                      // positions per thread and bulk-synchronous step
constexpr int LZF_HBITS = 13;
constexpr uint32_t LZF_TOK_EOS = 0x40000000u;     // literal-class token standing for the end-of-stream symbol

struct LzFrontSmem {
    __align__(16) uint16_t H[1 << LZF_HBITS];   // low 16 bits of (position + 1) of the most recent earlier position per hash (0 = none)
    uint32_t lenoff[LZF_TILE];         // (len-1) << 16 | offset; 0 = no match at this position
    uint32_t M[LZF_TILE / 32 + 1];     // match bitmap of the tile (+ a zero sentinel word)
    uint32_t V[LZF_TILE / 32];         // visited (= token start) bitmap of the tile
    uint32_t prefV[LZF_TILE / 32];     // exclusive prefix popcount of V
    uint32_t K[LZF_TILE / 32];         // visited matches (marks of the pointer-jumping parse)
    uint16_t nmw[LZF_TILE / 32 + 2];   // next non-zero word of M at or after w
    uint16_t pmw[LZF_TILE / 32];       // previous non-zero word of K before w (0xFFFF = none)
    __align__(16) uint16_t nxt[LZF_TILE];   // nxt(q) = first match position >= q + len[q] (relative; LZF_TILE = leaves the tile)
    uint16_t nxt2[LZF_TILE], nxt4[LZF_TILE];   // nxt o nxt and its square
    uint16_t chain[LZF_TILE / 12 + 8]; // every 4th visited match of the tile
    uint32_t nchain;
    uint32_t next_out;                 // furthest match end seen in this tile (relative)
    uint32_t next_p;                   // position (relative to the unit) where parsing resumes
    uint32_t ntok;                     // tokens emitted so far in the unit
};

// Word view of a unit for unaligned 4-byte reads: 32-bit index arithmetic on an aligned base instead of 64-bit pointer arithmetic
// per access.  last = index of the last word that holds a byte of the unit.
struct LzfSrc {
    const uint32_t* base;
    uint32_t skew, last;
    __device__ __forceinline__ LzfSrc(const uint8_t* src, uint32_t n)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        base = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
        skew = (uint32_t)(a & 3);
        last = n ? (skew + n - 1) >> 2 : 0u;
    }
};
// 4 bytes at unit offset p (little endian).  Only bytes below n are meaningful: every caller either masks the rest off (3-byte keys
// with p + 3 <= n) or stays below n (4-byte compare steps with i + 4 <= maxlen <= n - p).  Two aligned words and a funnel shift, no
// branch; the second word is clamped to the last word of the unit, so nothing beyond the unit's 4-byte granules is ever read.
__device__ __forceinline__ uint32_t lzf_ld4(const LzfSrc& S, uint32_t p)
{
    const uint32_t q = S.skew + p, i0 = q >> 2, i1 = min(i0 + 1u, S.last);
    return __funnelshift_r(__ldg(S.base + i0), __ldg(S.base + i1), 8u * (q & 3u));
}
__device__ __forceinline__ uint32_t lzf_hash3(uint32_t x) { return ((x & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - LZF_HBITS); }

// H[h] = low 16 bits of v = position + 1, unless the entry already holds a LATER position of the same step (step_base =
// first value of the step; its values span LZF_THREADS * LZF_STEP).  Entries of earlier steps are always older and are
// overwritten, so the result is the most recent position whatever the order of the threads: deterministic.
// 16-bit entries halve the table -- the kernels are latency-bound and the shared memory saved buys resident blocks.
__device__ __forceinline__ void lzf_insert(uint16_t* slot, uint32_t v, uint32_t step_base)
{
    const uint16_t nv = (uint16_t)v;
    if (nv == 0) return;                                           // 0 means "empty": such a position is simply not indexed
    const uint16_t mine = (uint16_t)(v - step_base);
    uint16_t old = *slot;
    for (;;) {
        const uint16_t theirs = (uint16_t)(old - (uint16_t)step_base);
        if (old != 0 && theirs < (uint16_t)(LZF_THREADS * LZF_STEP) && theirs >= mine) return;   // a later position of this step is in
        const uint16_t seen = atomicCAS(slot, old, nv);
        if (seen == old) return;
        old = seen;
    }
}

// common prefix length of src[p..] and src[c..] (c < p), at most maxlen
__device__ __forceinline__ uint32_t lzf_match_len(const LzfSrc& S, const uint8_t* __restrict__ src, uint32_t p, uint32_t c, uint32_t maxlen)
{
    uint32_t i = 0;
    while (i + 4 <= maxlen) {
        const uint32_t d = lzf_ld4(S, p + i) ^ lzf_ld4(S, c + i);
        if (d) return i + ((uint32_t)(__ffs((int)d) - 1) >> 3);
        i += 4;
    }
    while (i < maxlen && src[p + i] == src[c + i]) ++i;
    return i;
}

// Indexes positions [from, to) of src[0..n) without searching them: the history a segment-wise encoder wants in the table
// before it starts in the middle of a buffer.  Same insert order as lzf_tile.  Must be called by all LZF_THREADS threads.
__device__ __forceinline__ void lzf_warm(LzFrontSmem& sm, const uint8_t* __restrict__ src, uint32_t n, uint32_t from, uint32_t to)
{
    const int tid = (int)threadIdx.x;
    const LzfSrc S(src, n);
    for (uint32_t s0 = from; s0 < to; s0 += LZF_THREADS * LZF_STEP) {
#pragma unroll
        for (int u = 0; u < LZF_STEP; ++u) {
            const uint32_t p = s0 + (uint32_t)u * LZF_THREADS + tid;
            if (p < to && p + 3 <= n) lzf_insert(&sm.H[lzf_hash3(lzf_ld4(S, p) & 0xFFFFFFu)], p + 1, s0 + 1);
        }
        __syncthreads();
    }
}

// Greedy parse of tile [t0, t0 + LZF_TILE) of a unit of n bytes from sm.lenoff / sm.M (one entry / bit per position of the tile),
// entered at sm.next_p.  On return sm.V / sm.prefV describe the token starts and sm.next_p has advanced; the number of tokens
// starting in the tile is returned.  Must be called by all LZF_THREADS threads.
__device__ __forceinline__ uint32_t lzf_parse_tile(LzFrontSmem& sm, uint32_t n, uint32_t t0)
{
    const int tid = (int)threadIdx.x, lane = tid & 31;
    // Greedy parse of the tile, p += max(1, len[p]): for a match position q the parse continues at
    // nxt(q) = first match position >= q + len[q] (computed for all match positions in parallel); the visited matches
    // are the chain from the first match at or after the carried-in position -- one shared-memory load per link.
    {
        const uint32_t tend_r = min(t0 + LZF_TILE, n) - t0;                 // tile length (relative)
        const uint32_t start_r = sm.next_p > t0 ? sm.next_p - t0 : 0;       // >= tend_r when a long match skips the tile
        if (tid < 32) {     // nmw[w] = first word >= w of M that is not zero (LZF_TILE / 32 = none); 2 words per lane
            const int W = LZF_TILE / 32;
            const int w1 = 2 * lane + 1, w0 = 2 * lane;
            const int n1 = sm.M[w1] ? w1 : W;
            const int n0 = sm.M[w0] ? w0 : n1;
            int v = n0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_down_sync(FULL, v, d); if (lane + d < 32) v = min(v, t); }
            int after = __shfl_down_sync(FULL, v, 1);
            if (lane == 31) after = W;
            sm.nmw[w1] = (uint16_t)(n1 == W ? after : n1);
            sm.nmw[w0] = (uint16_t)(n0 == W ? after : n0);
            if (lane == 0) { sm.nmw[W] = (uint16_t)W; sm.M[W] = 0; }
        }
        if (tid >= 32 && tid < 32 + LZF_TILE / 32) sm.K[tid - 32] = 0;
        if (tid == 0) sm.next_out = 0;
        __syncthreads();
        auto first_match_from = [&](uint32_t e) -> uint32_t {               // relative position or LZF_TILE
            if (e >= (uint32_t)LZF_TILE) return LZF_TILE;
            uint32_t w = e >> 5;
            uint32_t bits = sm.M[w] & (0xffffffffu << (e & 31));
            if (!bits) { w = sm.nmw[w + 1]; bits = sm.M[w]; }                // M[LZF_TILE / 32] == 0
            return bits ? (w * 32 + (uint32_t)__ffs((int)bits) - 1) : (uint32_t)LZF_TILE;
        };
        uint16_t* cur = sm.nxt;
        for (int r = tid; r < LZF_TILE; r += LZF_THREADS)
            cur[r] = (uint16_t)(((sm.M[r >> 5] >> (r & 31)) & 1) ? first_match_from((uint32_t)r + (sm.lenoff[r] >> 16) + 1) : (uint32_t)LZF_TILE);
        __syncthreads();
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
        if (tid < 32) {     // pmw[w] = last word < w of K that is not zero (0xFFFF = none); 2 words per lane
            const int w0 = 2 * lane, w1 = 2 * lane + 1;
            const int k0 = sm.K[w0] ? w0 : -1;
            const int carry = sm.K[w1] ? w1 : k0;
            int v = carry;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, v, d); if (lane >= d) v = max(v, t); }
            int before = __shfl_up_sync(FULL, v, 1);
            if (lane == 0) before = -1;
            sm.pmw[w0] = (uint16_t)before;
            sm.pmw[w1] = (uint16_t)max(k0, before);
        }
        __syncthreads();
        // visited = a visited match, or a position at or after the carried-in start that the last visited match before it
        // does not cover (then it is a literal).  A warp handles one bitmap word per iteration: V needs no atomics.
        for (int r = tid; r < LZF_TILE; r += LZF_THREADS) {
            const int w = r >> 5;
            const uint32_t below = sm.K[w] & (0xffffffffu >> (31 - lane));
            int q = -1;
            if (below) q = w * 32 + 31 - __clz((int)below);
            else { const int w2 = (int)(int16_t)sm.pmw[w]; if (w2 >= 0) q = w2 * 32 + 31 - __clz((int)sm.K[w2]); }
            bool vis = (uint32_t)r >= start_r && (uint32_t)r < tend_r;
            if (q >= 0) {
                const uint32_t e = (uint32_t)q + (sm.lenoff[q] >> 16) + 1;
                vis = vis && (r == q || (uint32_t)r >= e);
                if (r == q) atomicMax(&sm.next_out, e);
            }
            const unsigned vv = __ballot_sync(FULL, vis);
            if (lane == 0) sm.V[w] = vv;
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


// Match search + greedy parse for tile [t0, t0 + LZF_TILE) of the unit src[0..n).  `window` = largest usable
// offset, `far3` = largest offset accepted for a 3-byte match.  On return sm.lenoff / sm.M / sm.V / sm.prefV
// describe the tile and sm.next_p has advanced; the number of tokens starting in the tile is returned.
// Must be called by all LZF_THREADS threads.
__device__ __forceinline__ uint32_t lzf_tile(LzFrontSmem& sm, const uint8_t* __restrict__ src, uint32_t n, uint32_t t0,
                                             uint32_t window, uint32_t far3)
{
    const int tid = (int)threadIdx.x, lane = tid & 31;
    const LzfSrc S(src, n);
    // LZF_STEP positions per thread and step: their lookups are independent, so every thread has that many chains of loads
    // in flight, and the two barriers of a step are paid once per LZF_STEP * 256 positions.
    for (uint32_t s0 = t0; s0 < t0 + LZF_TILE; s0 += LZF_THREADS * LZF_STEP) {
        uint32_t hh[LZF_STEP];
#pragma unroll
        for (int u = 0; u < LZF_STEP; ++u) {
            const uint32_t p = s0 + (uint32_t)u * LZF_THREADS + tid;
            uint32_t res = 0, h = 0xFFFFFFFFu;
            const bool on = p + 3 <= n;
            uint32_t key = 0;
            if (on) { key = lzf_ld4(S, p) & 0xFFFFFFu; h = lzf_hash3(key); }
            hh[u] = h;
            const unsigned peers = __match_any_sync(FULL, on ? h : (0x80000000u | (uint32_t)lane));
            if (on) {
                const uint32_t maxlen = min(n - p, LZQ_MAXLEN);
                uint32_t best = 0, boff = 0;
                const unsigned below = peers & ((1u << lane) - 1);
                if (below) {                                            // candidate B: nearest earlier lane, same hash
                    const uint32_t d = (uint32_t)(lane - (31 - __clz((int)below)));
                    if (d <= window && (lzf_ld4(S, p - d) & 0xFFFFFFu) == key) { best = lzf_match_len(S, src, p, p - d, maxlen); boff = d; }
                }
                const uint32_t a = sm.H[h];
                if (a) {                                                // candidate A: most recent position of earlier steps
                    // the table keeps 16 bits of the position: the distance is exact while it is below 65536; an older
                    // (aliased) entry yields some other earlier position, which the key compare accepts or rejects on its merits
                    const uint32_t d = (p + 1 - a) & 0xFFFFu, c = p - d;
                    if (d != 0 && d <= p && d <= window && d != boff && best < maxlen && (lzf_ld4(S, c) & 0xFFFFFFu) == key) {
                        const uint32_t l = lzf_match_len(S, src, p, c, maxlen);
                        if (l > best) { best = l; boff = d; }
                    }
                }
                if (best >= 3 && !(best == 3 && boff > far3)) res = ((best - 1) << 16) | boff;
            }
            sm.lenoff[p - t0] = res;
            const unsigned mm = __ballot_sync(FULL, res != 0);
            if (lane == 0) sm.M[(p - t0) >> 5] = mm;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < LZF_STEP; ++u)
            if (hh[u] != 0xFFFFFFFFu) lzf_insert(&sm.H[hh[u]], s0 + (uint32_t)u * LZF_THREADS + tid + 1, s0 + 1);
        __syncthreads();
    }
    return lzf_parse_tile(sm, n, t0);
}

} // namespace msc
