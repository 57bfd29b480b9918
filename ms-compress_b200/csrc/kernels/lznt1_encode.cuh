// lznt1_encode.cuh -- LZNT1 chunk encoder, one 256-thread block per 4 KiB chunk.
//
// Replaces the reference's lznt1_compress_chunk (src/lznt1_compress.cpp:49-94), its matcher
// LZNT1Dictionary::Fill/Find (include/mscomp/LZNT1Dictionary.h:93-106, :114-143) and the per-chunk part of
// lznt1_compress (:233-273: header, stored-chunk fallback).  Output is byte-identical to the reference:
// exhaustive longest match among all earlier positions with the same first three bytes, earliest position
// wins ties, stop at max_len; greedy parse; groups of 8 items behind a flag byte.
//
// Key fact: Find(p) is a pure function of (chunk, p), so it is evaluated for ALL positions in
// parallel; only the greedy walk p += max(1, len[p]) is serial, and it touches one word per match.
//
// Stages (all in shared memory):
//   A  stage the chunk, zero padded
//   B  stable radix sort of positions by a 12-bit hash of their first 3 bytes (ascending inside a bucket)
//   C  all-position match search -> lenoff[p] = len | dist << 16, match bitmap M
//   D  greedy parse: one thread walks M with find-first-set and records the visited bitmap V
//   E  every visited position computes its item index k and the number of matches before it from prefix
//      popcounts of V and V&M:  byte offset = k + (k >> 3) + 1 + matches_before;  flag bits by atomicOr
//   F  header + payload to the chunk's slot (compaction to a contiguous stream is a separate kernel)
#pragma once
#include "common.cuh"
#include "lznt1_decode.cuh"   // lznt1_shift_limit, LZNT1_CHUNK

namespace msc {

constexpr int LZE_THREADS = 256;
constexpr int LZE_BUCKETS = 4096;
constexpr int LZNT1_SLOT = 4112;                 // 2-byte header + <= 4096 payload, rounded up to 16

struct Lznt1EncSmem {
    __align__(16) uint8_t in[LZNT1_CHUNK + 16];  // zero padded
    union {
        uint32_t cnt[LZE_BUCKETS];               // B,C: bucket counts -> end offsets
        uint8_t out[LZNT1_SLOT];                 // E,F: assembled chunk (aliases cnt, which is dead by then)
    };
    uint16_t sorted[LZNT1_CHUNK];                // positions grouped by hash, ascending inside a group
    uint16_t ord[LZNT1_CHUNK];                   // indices into sorted[], grouped by rank inside their bucket
    uint16_t hashv[LZNT1_CHUNK];                 // 12-bit hash of the 3 bytes at every indexed position
    uint32_t rankcnt[256];                       // entries per rank (ranks >= 255 share the last bin)
    uint32_t lenoff[LZNT1_CHUNK];                // len | dist << 16 (0 = no match)
    uint32_t M[128], V[128], F[128];             // match bitmap, visited bitmap (by position), flag bits (by item)
    uint16_t prefV[128], prefVM[128];
    uint16_t goff[LZNT1_CHUNK / 8 + 8];          // byte offset of each group's flag byte
    uint16_t chain[LZNT1_CHUNK / 12 + 8];        // every 4th visited match of the greedy parse
    uint32_t nchain, scan_tmp[8];
};

__device__ __forceinline__ uint32_t lz_load4(const uint8_t* s, int idx)
{   // unaligned 4-byte read from a 4-byte aligned shared array
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s) + (idx >> 2);
    return __funnelshift_r(w[0], w[1], 8 * (idx & 3));
}
__device__ __forceinline__ uint32_t lz_hash3(uint32_t x) { return ((x & 0xFFFFFFu) * 0x9E3779B1u) >> 20; }

__global__ void __launch_bounds__(LZE_THREADS)
lznt1_encode_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t first, uint32_t nchunks,
                    uint8_t* __restrict__ slots, uint32_t* __restrict__ slot_size)
{
    MSC_DYN_SMEM(smem_raw);
    Lznt1EncSmem& sm = *reinterpret_cast<Lznt1EncSmem*>(smem_raw);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // chunks [first, first + nchunks) of the stream `in` (a slice, so that uploads can overlap with encoding)
    for (uint32_t c = first + blockIdx.x; c < first + nchunks; c += gridDim.x) {
        const uint64_t base = (uint64_t)c * LZNT1_CHUNK;
        const int n = (int)((in_len - base < (uint64_t)LZNT1_CHUNK) ? (in_len - base) : LZNT1_CHUNK);
        const uint8_t* src = in + base;
        uint8_t* slot = slots + (uint64_t)c * LZNT1_SLOT;

        // ---- A: stage ------------------------------------------------------------------------------------
        if ((((uintptr_t)src) & 15u) == 0 && n == LZNT1_CHUNK) {
            reinterpret_cast<uint4*>(sm.in)[tid] = ld_stream16(src + 16 * tid);
        } else {
            for (int i = tid; i < LZNT1_CHUNK; i += LZE_THREADS) sm.in[i] = (i < n) ? src[i] : 0;
        }
        if (tid < 16) sm.in[LZNT1_CHUNK + tid] = 0;
        for (int i = tid; i < LZE_BUCKETS; i += LZE_THREADS) sm.cnt[i] = 0;
        if (tid < 128) { sm.V[tid] = 0; sm.F[tid] = 0; }
        __syncthreads();

        // ---- B: counting sort by hash ------------------------------------------------------------------
        const int npos = n - 2;                      // positions 0..n-3 are indexed (Fill :99)
        for (int p = tid; p < npos; p += LZE_THREADS) {
            const uint32_t h = lz_hash3(lz_load4(sm.in, p));
            sm.hashv[p] = (uint16_t)h;
            atomicAdd(&sm.cnt[h], 1u);
        }
        __syncthreads();
        {   // inclusive scan of 4096 counts (16 per thread + block scan of the 256 partial sums): cnt[h] = END of bucket h
            uint32_t loc[16], sum = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { loc[i] = sm.cnt[tid * 16 + i]; sum += loc[i]; }
            int incl = warp_incl_scan((int)sum);
            if (lane == 31) sm.scan_tmp[warp] = (uint32_t)incl;
            __syncthreads();
            uint32_t woff = 0;
            for (int w = 0; w < warp; ++w) woff += sm.scan_tmp[w];
            uint32_t run = woff + (uint32_t)incl - sum;
#pragma unroll
            for (int i = 0; i < 16; ++i) { run += loc[i]; sm.cnt[tid * 16 + i] = run; }
        }
        // Stable LSD radix sort of the positions by their 12-bit hash (3 passes of 4 bits): every thread owns 16
        // consecutive elements and a private column of 16 digit counters, so ranks need no atomics and equal keys
        // keep their order -- positions stay ascending inside a bucket, which the early exit of the search relies on.
        {
            uint16_t* cntT = reinterpret_cast<uint16_t*>(sm.lenoff);          // [16 digits][256 threads], lenoff is not live yet
            const uint16_t* srcA = sm.ord;                                    // pass 0 reads the identity instead
            uint16_t* dstA = sm.sorted;
            for (int pass = 0; pass < 3; ++pass) {
                const int sh = 4 * pass;
#pragma unroll
                for (int d = 0; d < 16; ++d) cntT[d * LZE_THREADS + tid] = 0;
                for (int i = 0; i < 16; ++i) {
                    const int e = tid * 16 + i;
                    if (e < npos) {
                        const int p = pass ? (int)srcA[e] : e;
                        cntT[(((uint32_t)sm.hashv[p] >> sh) & 15) * LZE_THREADS + tid] += 1;
                    }
                }
                __syncthreads();
                {   // exclusive scan of the 4096 counters in (digit, thread) order
                    uint32_t loc[16], sum = 0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { loc[i] = cntT[tid * 16 + i]; sum += loc[i]; }
                    const int incl = warp_incl_scan((int)sum);
                    if (lane == 31) sm.scan_tmp[warp] = (uint32_t)incl;
                    __syncthreads();
                    uint32_t woff = 0;
                    for (int w = 0; w < warp; ++w) woff += sm.scan_tmp[w];
                    uint32_t run = woff + (uint32_t)incl - sum;
#pragma unroll
                    for (int i = 0; i < 16; ++i) { cntT[tid * 16 + i] = (uint16_t)run; run += loc[i]; }
                }
                __syncthreads();
                for (int i = 0; i < 16; ++i) {
                    const int e = tid * 16 + i;
                    if (e < npos) {
                        const int p = pass ? (int)srcA[e] : e;
                        const int c = (((uint32_t)sm.hashv[p] >> sh) & 15) * LZE_THREADS + tid;
                        const uint32_t slot = cntT[c];
                        cntT[c] = (uint16_t)(slot + 1);
                        dstA[slot] = (uint16_t)p;
                    }
                }
                __syncthreads();
                const uint16_t* t2 = dstA; dstA = const_cast<uint16_t*>(srcA); srcA = t2;     // sorted -> ord -> sorted
            }
        }
        // cnt[h] = end offset of bucket h; begin = cnt[h-1] (0 for h == 0); sorted[] holds the result of pass 2

        // ---- C: match search for every position ------------------------------------------------------
        // The k-th entry of a bucket has k candidates (the earlier entries).  Threads take the entries grouped by that
        // rank, so the lanes of a warp run the candidate loop the same number of times.
        for (int p = (npos > 0 ? npos : 0) + tid; p < LZNT1_CHUNK; p += LZE_THREADS) sm.lenoff[p] = 0;   // not indexed: no match
        sm.rankcnt[tid] = 0;
        __syncthreads();
        for (int idx = tid; idx < npos; idx += LZE_THREADS) {
            const uint32_t h = sm.hashv[sm.sorted[idx]];
            const int r = idx - (int)(h ? sm.cnt[h - 1] : 0);
            atomicAdd(&sm.rankcnt[r < 255 ? r : 255], 1u);
        }
        __syncthreads();
        {   // exclusive scan of the 256 rank counts (one per thread)
            const uint32_t v = sm.rankcnt[tid];
            const int incl = warp_incl_scan((int)v);
            if (lane == 31) sm.scan_tmp[warp] = (uint32_t)incl;
            __syncthreads();
            uint32_t woff = 0;
            for (int w = 0; w < warp; ++w) woff += sm.scan_tmp[w];
            sm.rankcnt[tid] = woff + (uint32_t)incl - v;
        }
        __syncthreads();
        for (int idx = tid; idx < npos; idx += LZE_THREADS) {
            const uint32_t h = sm.hashv[sm.sorted[idx]];
            const int r = idx - (int)(h ? sm.cnt[h - 1] : 0);
            sm.ord[atomicAdd(&sm.rankcnt[r < 255 ? r : 255], 1u)] = (uint16_t)idx;
        }
        __syncthreads();
        for (int i = tid; i < npos; i += LZE_THREADS) {
            const int idx = sm.ord[i];
            const int p = sm.sorted[idx];
            uint32_t res = 0;
            if (p > 0) {
                int S, lim;
                lznt1_shift_limit(p, S, lim);
                const int max_len = min(n - p, (1 << S) + 2);       // mask3 (src/lznt1_compress.cpp:51,:66)
                const uint32_t key = lz_load4(sm.in, p);
                const uint32_t h = sm.hashv[p];
                int k = (h == 0) ? 0 : (int)sm.cnt[h - 1];
                int best = 0, bj = 0;
                for (; k < idx; ++k) {                              // earlier entries of my bucket = earlier positions
                    const int j = sm.sorted[k];
                    // common prefix of p and j, 4 bytes at a time; a hash collision simply ends with i < 3
                    int i = 0;
                    uint32_t d = lz_load4(sm.in, j) ^ key;
                    while (!d) {
                        i += 4;
                        if (i >= max_len) break;
                        d = lz_load4(sm.in, p + i) ^ lz_load4(sm.in, j + i);
                    }
                    if (d) i += (__ffs((int)d) - 1) >> 3;
                    if (i > max_len) i = max_len;
                    if (i > best) { best = i; bj = j; if (best == max_len) break; }
                }
                if (best >= 3) res = (uint32_t)best | ((uint32_t)(p - bj) << 16);
            }
            sm.lenoff[p] = res;
        }
        __syncthreads();
        for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) {
            const unsigned mm = __ballot_sync(FULL, sm.lenoff[p] != 0);
            if (lane == 0) sm.M[p >> 5] = mm;
        }
        __syncthreads();

        // ---- D: greedy parse p += max(1, len[p]) ------------------------------------------
        // For a match position q the parse continues at nxt(q) = first match position >= q + len[q] (everything in
        // between is literals); nxt is computed for all match positions in parallel, the visited ones are then the
        // chain from the first match -- one shared-memory load per visited match.
        {
            uint16_t* cur = sm.sorted;                                                        // dead after stage C
            auto first_match_from = [&](int e) -> int {
                if (e >= LZNT1_CHUNK) return LZNT1_CHUNK;
                int w = e >> 5;
                uint32_t bits = sm.M[w] & (0xffffffffu << (e & 31));
                while (!bits && ++w < 128) bits = sm.M[w];
                return bits ? (w * 32 + __ffs((int)bits) - 1) : LZNT1_CHUNK;
            };
            auto or_range = [&](int lo, int hi) {                                             // V |= [lo, hi)
                for (int w = lo >> 5; lo < hi && w <= ((hi - 1) >> 5); ++w) {
                    const int lo_b = max(lo, w * 32) - w * 32, hi_b = min(hi, w * 32 + 32) - w * 32;
                    atomicOr(&sm.V[w], (hi_b == 32 ? 0xffffffffu : ((1u << hi_b) - 1)) & (0xffffffffu << lo_b));
                }
            };
            uint16_t* nxt2 = sm.ord;                                                          // dead after stage C as well
            uint16_t* nxt4 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(sm.cnt) + 8192);   // upper half of cnt: dead, clear of `out`
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS)
                cur[p] = (uint16_t)(((sm.M[p >> 5] >> (p & 31)) & 1) ? first_match_from(p + (int)(sm.lenoff[p] & 0xFFFF)) : LZNT1_CHUNK);
            __syncthreads();                                                                  // F doubles as the mark bitmap
            if (tid == 32) or_range(0, min(first_match_from(0), n));                          // leading literals
            // two rounds of pointer doubling shorten the serial part of the chase by 4: nxt2 = nxt o nxt, nxt4 = nxt2 o nxt2
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) { const int t = cur[p]; nxt2[p] = (uint16_t)(t < LZNT1_CHUNK ? cur[t] : LZNT1_CHUNK); }
            __syncthreads();
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) { const int t = nxt2[p]; nxt4[p] = (uint16_t)(t < LZNT1_CHUNK ? nxt2[t] : LZNT1_CHUNK); }
            __syncthreads();
            if (tid == 0) {     // every 4th visited match: one shared-memory load per link
                int q = first_match_from(0), k = 0;
                while (q < LZNT1_CHUNK) { sm.chain[k++] = (uint16_t)q; q = nxt4[q]; }
                sm.nchain = (uint32_t)k;
            }
            __syncthreads();
            for (uint32_t k = tid; k < sm.nchain; k += LZE_THREADS) {                         // the three in between, in parallel
                int q = sm.chain[k];
                atomicOr(&sm.F[q >> 5], 1u << (q & 31));
                const int q2 = nxt2[q];
                q = cur[q];
                if (q < LZNT1_CHUNK) atomicOr(&sm.F[q >> 5], 1u << (q & 31));
                if (q2 < LZNT1_CHUNK) { atomicOr(&sm.F[q2 >> 5], 1u << (q2 & 31)); const int q3 = cur[q2]; if (q3 < LZNT1_CHUNK) atomicOr(&sm.F[q3 >> 5], 1u << (q3 & 31)); }
            }
            __syncthreads();
            // visited bitmap: every visited match plus the literals up to the next visited match
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) {
                if ((sm.F[p >> 5] >> (p & 31)) & 1) {
                    const int e = p + (int)(sm.lenoff[p] & 0xFFFF);
                    or_range(p, p + 1);
                    or_range(e, min(first_match_from(e), n));
                }
            }
            __syncthreads();
            if (tid < 128) sm.F[tid] = 0;
        }
        __syncthreads();

        // ---- E: offsets and emission ------------------------------------------------------------------
        if (warp == 0) {   // exclusive prefix popcounts over the 128 bitmap words (4 per lane)
            uint32_t a[4], b[4], sa = 0, sb = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t v = sm.V[lane * 4 + i];
                a[i] = (uint32_t)__popc(v); b[i] = (uint32_t)__popc(v & sm.M[lane * 4 + i]);
                sa += a[i]; sb += b[i];
            }
            uint32_t ra = (uint32_t)warp_incl_scan((int)sa) - sa, rb = (uint32_t)warp_incl_scan((int)sb) - sb;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                sm.prefV[lane * 4 + i] = (uint16_t)ra; sm.prefVM[lane * 4 + i] = (uint16_t)rb;
                ra += a[i]; rb += b[i];
            }
        }
        __syncthreads();
        const uint32_t items = sm.prefV[127] + (uint32_t)__popc(sm.V[127]);
        const uint32_t matches = sm.prefVM[127] + (uint32_t)__popc(sm.V[127] & sm.M[127]);
        const uint32_t ngroups = (items + 7) >> 3;
        const uint32_t total = items + ngroups + matches;
        const bool compressed = total < (uint32_t)n;                 // src/lznt1_compress.cpp:86, :245
        if (compressed) {
            uint8_t* o = sm.out + 2;
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) {
                const uint32_t v = sm.V[p >> 5], below = (1u << (p & 31)) - 1;
                if ((v >> (p & 31)) & 1) {
                    const uint32_t vm = v & sm.M[p >> 5];
                    const uint32_t k = sm.prefV[p >> 5] + (uint32_t)__popc(v & below);
                    const uint32_t m = sm.prefVM[p >> 5] + (uint32_t)__popc(vm & below);
                    const uint32_t off = k + (k >> 3) + 1 + m;
                    if ((k & 7) == 0) sm.goff[k >> 3] = (uint16_t)(off - 1);
                    if ((vm >> (p & 31)) & 1) {
                        int S, lim;
                        lznt1_shift_limit(p, S, lim);
                        const uint32_t lo = sm.lenoff[p];
                        const uint32_t tok = (((lo >> 16) - 1) << S) | ((lo & 0xFFFF) - 3);
                        o[off] = (uint8_t)tok; o[off + 1] = (uint8_t)(tok >> 8);
                        atomicOr(&sm.F[k >> 5], 1u << (k & 31));
                    } else {
                        o[off] = sm.in[p];
                    }
                }
            }
            __syncthreads();
            for (uint32_t g = tid; g < ngroups; g += LZE_THREADS)
                o[sm.goff[g]] = (uint8_t)(sm.F[g >> 2] >> (8 * (g & 3)));
            if (tid == 0) { const uint32_t hdr = 0xB000u | (total - 1); sm.out[0] = (uint8_t)hdr; sm.out[1] = (uint8_t)(hdr >> 8); }
        } else {
            for (int i = tid; i < n; i += LZE_THREADS) sm.out[2 + i] = sm.in[i];
            if (tid == 0) { const uint32_t hdr = 0x3000u | (uint32_t)(n - 1); sm.out[0] = (uint8_t)hdr; sm.out[1] = (uint8_t)(hdr >> 8); }
        }
        __syncthreads();

        // ---- F: slot write ---------------------------------------------------------------------------------
        const uint32_t bytes = 2 + (compressed ? total : (uint32_t)n);
        for (uint32_t v = tid; v * 16 < bytes; v += LZE_THREADS)
            st_stream16(slot + 16 * v, *reinterpret_cast<const uint4*>(sm.out + 16 * v));
        if (tid == 0) slot_size[c] = bytes;
        __syncthreads();
    }
}

} // namespace msc
