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
// Stages (all in shared memory, 50 KB per block -> 4 blocks per SM):
//   A  stage the chunk, zero padded
//   B  12-bit hash of the 3 bytes at every position; histogram -> bucket begin offsets; ONE warp then places the
//      positions in position order, 32 per round (__match_any_sync gives the rank among equal hashes of the round), so
//      buckets come out ascending by position -- a stable counting sort without per-thread counter columns
//   C  all-position match search -> lenoff[p] = len | dist << 16, match bitmap M; entries are handed to threads
//      grouped by their rank inside the bucket (= number of candidates), rank-0 entries are skipped
//   D  greedy parse: successor links between match positions, doubled twice, one thread chases every 4th visited
//      match, the rest is filled in parallel; visited bitmap V by ballot
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
// A chunk whose all-position search would cost more than LZE_HEAVY_WORK candidate evaluations (long runs, tiny alphabets:
// huge buckets) is "heavy": entries of rank >= LZE_LAZY_RANK are then searched only if the greedy parse visits them
// (as the reference does for every position, src/lznt1_compress.cpp:62-70), by a whole warp per visited position.
constexpr uint32_t LZE_HEAVY_WORK = 32768;
constexpr int LZE_LAZY_RANK = 16;
constexpr int LZE_ITEM_CAP = LZNT1_CHUNK / 2;    // work items of the pair search (they live in ord[], 4 bytes each)

struct Lznt1EncSmem {
    __align__(16) uint8_t in[LZNT1_CHUNK + 16];  // zero padded; a full, aligned chunk arrives by one bulk async copy (cp.async.bulk + mbarrier)
    __align__(8) uint64_t bar;                   // its mbarrier
    union {
        __align__(16) uint16_t cnt[LZE_BUCKETS]; // B,C: bucket cursors: begin offsets -> (after placement) end offsets
        uint16_t nxt4[LZNT1_CHUNK];              // D: successor link doubled twice
        uint8_t out[LZNT1_SLOT];                 // E,F: assembled chunk
    };
    __align__(16) uint16_t sorted[LZNT1_CHUNK];  // B,C: positions grouped by hash, ascending inside a group; D: nxt
    __align__(16) uint16_t ord[LZNT1_CHUNK];     // C: indices into sorted[], grouped by rank inside their bucket; D: nxt2
    __align__(16) uint32_t lenoff[LZNT1_CHUNK];  // C..E: len | dist << 16 (0 = no match); B: its first half holds the hashes
    uint32_t rankcnt[256];                       // entries per rank (ranks >= 255 share the last bin)
    uint32_t M[132], V[128], F[128], K[128];     // match bitmap (+ zero sentinel), visited bitmap, flag bits (by item), visited matches
    uint32_t U[128];                             // heavy chunks: positions whose search is deferred to the parse
    uint32_t C[128];                             // positions inside (not at the start of) a match the parse takes
    uint16_t nmw[132], pmw[128];                 // next non-zero word of M at or after w; previous non-zero word of K before w
    uint16_t prefV[128], prefVM[128];
    uint16_t goff[LZNT1_CHUNK / 8 + 8];          // byte offset of each group's flag byte
    uint16_t chain[LZNT1_CHUNK / 12 + 8];        // every 4th visited match of the greedy parse
    uint32_t nchain, scan_tmp[8], work_tmp[8], cross_tmp[8], nbig, big_fill[4];
    unsigned long long tot_tmp[8], small_tmp[8];
};

__device__ __forceinline__ uint32_t lz_load4(const uint8_t* s, int idx)
{   // unaligned 4-byte read from a 4-byte aligned shared array
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s) + (idx >> 2);
    return __funnelshift_r(w[0], w[1], 8 * (idx & 3));
}
__device__ __forceinline__ uint32_t lz_hash3(uint32_t x) { return ((x & 0xFFFFFFu) * 0x9E3779B1u) >> 20; }

// A match the greedy parse takes at q covers positions q+1 .. q+len-1: they start no item.
__device__ __forceinline__ void lze_cover(uint32_t* C, int q, int len)
{
    const int s = q + 1, e = q + len - 1;                                     // (len >= 3)
    const int w0 = s >> 5, w1 = e >> 5;
    const uint32_t m0 = 0xffffffffu << (s & 31), m1 = 0xffffffffu >> (31 - (e & 31));
    if (w0 == w1) atomicOr(&C[w0], m0 & m1);
    else { atomicOr(&C[w0], m0); for (int w = w0 + 1; w < w1; ++w) atomicOr(&C[w], 0xffffffffu); atomicOr(&C[w1], m1); }
}

// One position of a bucket in registers: its 16 bytes, the longest match the format can express there, the best key so far.
struct LzeEnt { int p, room; uint32_t k0, k1, k2, k3, best; };          // room = bytes from p to the end of the chunk
__device__ __forceinline__ void lze_load(LzeEnt& e, const Lznt1EncSmem& sm, int idx, int n)
{
    e.p = sm.sorted[idx];
    const uint32_t* w = reinterpret_cast<const uint32_t*>(sm.in) + (e.p >> 2);
    const uint32_t sh = 8u * ((uint32_t)e.p & 3u);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];     // (in[] is zero padded by 16 bytes)
    e.k0 = __funnelshift_r(w0, w1, sh); e.k1 = __funnelshift_r(w1, w2, sh); e.k2 = __funnelshift_r(w2, w3, sh); e.k3 = __funnelshift_r(w3, w4, sh);
    e.room = n - e.p;
    e.best = 0;
}
// Common prefix of two positions of a bucket (16 bytes without a branch, the rare rest from shared memory), credited to the later one.
__device__ __forceinline__ void lze_pair(LzeEnt& x, LzeEnt& y, const Lznt1EncSmem& sm)
{
    uint32_t w = x.k0 ^ y.k0;
    int l = 0;
    if (!w) { w = x.k1 ^ y.k1; l = 4; }
    if (!w) { w = x.k2 ^ y.k2; l = 8; }
    if (!w) { w = x.k3 ^ y.k3; l = 12; }
    l = w ? l + ((__ffs((int)w) - 1) >> 3) : 16;
    const bool x_later = x.p > y.p;
    const int room = x_later ? x.room : y.room;
    // max_len(p) = min(room, (1 << S(p)) + 2) (mask3, src/lznt1_compress.cpp:51,:66) is at least min(room, 18): the format's
    // cap only matters once 16 bytes have matched
    if (l == 16 && room > 16) {
        int S, lim;
        lznt1_shift_limit(x_later ? x.p : y.p, S, lim);
        const int ml = min(room, (1 << S) + 2);
        while (l < ml) {
            const uint32_t d = lz_load4(sm.in, x.p + l) ^ lz_load4(sm.in, y.p + l);
            if (d) { l += (__ffs((int)d) - 1) >> 3; break; }
            l += 4;
        }
        l = min(l, ml);
    }
    l = min(l, room);
    const uint32_t key = l >= 3 ? ((uint32_t)l << 12) | (uint32_t)(4095 - (x_later ? y.p : x.p)) : 0u;
    x.best = max(x.best, x_later ? key : 0u);
    y.best = max(y.best, x_later ? 0u : key);
}

__global__ void __launch_bounds__(LZE_THREADS, 4)
lznt1_encode_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t first, uint32_t nchunks,
                    uint8_t* __restrict__ slots, uint32_t* __restrict__ slot_size, uint32_t heavy_work)
{
    MSC_DYN_SMEM(smem_raw);
    Lznt1EncSmem& sm = *reinterpret_cast<Lznt1EncSmem*>(smem_raw);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t* hashv = reinterpret_cast<uint16_t*>(sm.lenoff);           // live in B and in the rank pass of C only

    if (tid == 0) mbar_init(&sm.bar, 1);
    __syncthreads();
    uint32_t phase = 0;
    // chunks [first, first + nchunks) of the stream `in` (a slice, so that uploads can overlap with encoding)
    for (uint32_t c = first + blockIdx.x; c < first + nchunks; c += gridDim.x) {
        const uint64_t base = (uint64_t)c * LZNT1_CHUNK;
        const int n = (int)((in_len - base < (uint64_t)LZNT1_CHUNK) ? (in_len - base) : LZNT1_CHUNK);
        const uint8_t* src = in + base;
        uint8_t* slot = slots + (uint64_t)c * LZNT1_SLOT;

        // ---- A: stage ------------------------------------------------------------------------------------
        const bool bulk = (((uintptr_t)src) & 15u) == 0 && n == LZNT1_CHUNK;      // (block-uniform)
        if (bulk) {
            if (tid == 0) { fence_async_smem(); bulk_load_g2s(sm.in, src, LZNT1_CHUNK, &sm.bar); }   // the chunk: one bulk copy, global -> shared
        } else {
            for (int i = tid; i < LZNT1_CHUNK; i += LZE_THREADS) sm.in[i] = (i < n) ? src[i] : 0;
        }
        if (tid < 16) sm.in[LZNT1_CHUNK + tid] = 0;
        reinterpret_cast<uint4*>(sm.cnt)[tid] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(sm.cnt)[tid + LZE_THREADS] = make_uint4(0, 0, 0, 0);
        if (tid < 128) { sm.F[tid] = 0; sm.C[tid] = 0; sm.U[tid] = 0; }
        sm.rankcnt[tid] = 0;
        if (bulk) { mbar_wait(&sm.bar, phase); phase ^= 1u; }
        __syncthreads();

        // ---- B: hash of every position + histogram ------------------------------------------------------
        const int npos = n - 2;                      // positions 0..n-3 are indexed (Fill :99)
        uint32_t hh[8];                              // my 16 hashes, two per register (0xFFFF: no position)
        {
            uint32_t* cntw = reinterpret_cast<uint32_t*>(sm.cnt);     // two 16-bit counters per word (counts <= 4094)
#pragma unroll
            for (int m = 0; m < LZNT1_CHUNK / LZE_THREADS; ++m) {
                const int p = tid + LZE_THREADS * m;
                uint32_t h = 0xFFFFu;
                if (p < npos) {
                    h = lz_hash3(lz_load4(sm.in, p));
                    hashv[p] = (uint16_t)h;
                    atomicAdd(&cntw[h >> 1], 1u << (16 * (h & 1)));
                }
                if (m & 1) hh[m >> 1] |= h << 16; else hh[m >> 1] = h;
            }
        }
        __syncthreads();
        // Exclusive scan of the 4096 counts (16 per thread + block scan of the 256 partial sums): cnt[h] = BEGIN of bucket h.
        // The same pass lists the work of the search.  A bucket of c positions is cut into tiles of 4; every pair of positions of a
        // bucket is compared exactly once, inside a tile ("in-tile" item; lists by tile size 2, 3, 4 so that a warp's lanes do the
        // same work) or between two tiles ("cross" item).  Buckets of up to 4 are one tile, listed right here; the few larger ones
        // go to a side list and are expanded below.  Item lists live in ord[]: [size 2 | size 3 | size 4 | cross].
        bool fast;                                   // block-uniform: the pair search below; else the ordered-bucket path (heavy chunks)
        uint32_t base3 = 0, base4 = 0, basex = 0, n_items = 0;
        {
            uint4 q0 = reinterpret_cast<const uint4*>(sm.cnt)[2 * tid], q1 = reinterpret_cast<const uint4*>(sm.cnt)[2 * tid + 1];
            uint32_t wv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t sum = 0, my_pairs = 0, my_big = 0, bigmask = 0, s2 = 0, s3 = 0, s4 = 0;
            unsigned long long tot = 0;              // tiles of size 2 | 3 << 16 | 4 << 32 | cross items << 48, large buckets only
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t c = (wv[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
                sum += c;
                my_pairs += (c * (c - 1)) >> 1;
                s2 += c == 2 ? 1u : 0u; s3 += c == 3 ? 1u : 0u; s4 += c == 4 ? 1u : 0u;
                bigmask |= c > 4 ? 1u << i : 0u;
            }
            for (uint32_t bm = bigmask; bm; bm &= bm - 1) {                  // the few large buckets (counts still in cnt[])
                const uint32_t c = sm.cnt[16 * tid + (__ffs((int)bm) - 1)];
                const uint32_t T = (c + 3) >> 2, sl = c - 4 * (T - 1);
                ++my_big;
                tot += ((unsigned long long)(T - 1) << 32) + ((unsigned long long)((T * (T - 1)) >> 1) << 48) + (sl >= 2 ? 1ull << (16 * (sl - 2)) : 0ull);
            }
            const unsigned long long small = (unsigned long long)s2 | ((unsigned long long)s3 << 16) | ((unsigned long long)s4 << 32);   // single-tile buckets by size
            const int incl = warp_incl_scan((int)sum);
            unsigned long long incl_small = small;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, incl_small, d); if (lane >= d) incl_small += t; }
            const uint32_t wp = __reduce_add_sync(FULL, my_pairs), wb = __reduce_add_sync(FULL, my_big);
            unsigned long long wt = tot;
#pragma unroll
            for (int d = 16; d; d >>= 1) wt += __shfl_xor_sync(FULL, wt, d);
            if (lane == 31) { sm.scan_tmp[warp] = (uint32_t)incl; sm.work_tmp[warp] = wp; sm.small_tmp[warp] = incl_small; sm.cross_tmp[warp] = wb; sm.tot_tmp[warp] = wt; }
            if (tid == 0) { sm.nbig = 0; sm.big_fill[0] = 0; sm.big_fill[1] = 0; sm.big_fill[2] = 0; sm.big_fill[3] = 0; }
            __syncthreads();
            uint32_t woff = 0, pairs = 0, nbig = 0;
            unsigned long long big_tot = 0, soff = 0, nsmall = 0;
            for (int w = 0; w < LZE_THREADS / 32; ++w) {
                if (w < warp) { woff += sm.scan_tmp[w]; soff += sm.small_tmp[w]; }
                pairs += sm.work_tmp[w]; nsmall += sm.small_tmp[w]; nbig += sm.cross_tmp[w]; big_tot += sm.tot_tmp[w];
            }
            const uint32_t n2s = (uint32_t)(nsmall & 0xFFFFu), n3s = (uint32_t)((nsmall >> 16) & 0xFFFFu), n4s = (uint32_t)(nsmall >> 32);
            const uint32_t n2 = n2s + (uint32_t)(big_tot & 0xFFFFu), n3 = n3s + (uint32_t)((big_tot >> 16) & 0xFFFFu), n4 = n4s + (uint32_t)((big_tot >> 32) & 0xFFFFu);
            const uint32_t nx = (uint32_t)(big_tot >> 48);
            base3 = n2; base4 = n2 + n3; basex = n2 + n3 + n4; n_items = basex + nx;
            // (16-bit fields: a chunk whose counts would overflow them has far more pairs than heavy_work)
            fast = pairs <= heavy_work && n_items <= (uint32_t)LZE_ITEM_CAP && nbig <= 256u;
            soff += incl_small - small;
            uint32_t o2 = (uint32_t)(soff & 0xFFFFu), o3 = base3 + (uint32_t)((soff >> 16) & 0xFFFFu), o4 = base4 + (uint32_t)(soff >> 32);
            uint32_t run = woff + (uint32_t)incl - sum;
            uint32_t* items = reinterpret_cast<uint32_t*>(sm.ord);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t c = (wv[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
                const uint32_t b = run;
                run += c;
                if (i & 1) wv[i >> 1] = (wv[i >> 1] & 0xFFFFu) | (b << 16); else wv[i >> 1] = (wv[i >> 1] & 0xFFFF0000u) | b;
                if (fast) {
                    if (c == 2) items[o2++] = b;
                    if (c == 3) items[o3++] = b;
                    if (c == 4) items[o4++] = b;
                }
            }
            const uint32_t run_end = run;
            reinterpret_cast<uint4*>(sm.cnt)[2 * tid] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            reinterpret_cast<uint4*>(sm.cnt)[2 * tid + 1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
            if (fast) {
                for (uint32_t bm = bigmask; bm; bm &= bm - 1) {              // (my own stores to cnt[] above: begin offsets by now)
                    const int i = __ffs((int)bm) - 1;
                    const uint32_t b = sm.cnt[16 * tid + i], e = i < 15 ? (uint32_t)sm.cnt[16 * tid + i + 1] : run_end;
                    sm.rankcnt[atomicAdd(&sm.nbig, 1u)] = b | ((e - b) << 12);    // (rankcnt[] is free on this path)
                }
            }
            __syncthreads();
            if (fast && (uint32_t)tid < nbig) {       // a large bucket: its tiles and tile pairs (places by atomics: the order inside a list is free)
                const uint32_t e = sm.rankcnt[tid], b = e & 0xFFFu, c = e >> 12;
                const uint32_t T = (c + 3) >> 2, sl = c - 4 * (T - 1);
                uint32_t at = base4 + n4s + atomicAdd(&sm.big_fill[2], T - 1);
                for (uint32_t a = 0; a + 1 < T; ++a) items[at++] = b + 4 * a;
                if (sl >= 2) {
                    const uint32_t cls = sl - 2, basec = cls == 0 ? n2s : (cls == 1 ? base3 + n3s : base4 + n4s);
                    items[basec + atomicAdd(&sm.big_fill[cls], 1u)] = b + 4 * (T - 1);
                }
                at = basex + atomicAdd(&sm.big_fill[3], (T * (T - 1)) >> 1);
                for (uint32_t a = 1; a < T; ++a) {
                    const uint32_t A = (b + 4 * a) | (((a + 1 < T ? 4u : sl) - 1) << 12);
                    for (uint32_t bb = 0; bb < a; ++bb) items[at++] = A | ((b + 4 * bb) << 14);          // (earlier tiles are full)
                }
            }
        }
        __syncthreads();

        bool heavy = false;                          // block-uniform: deferred (lazy) search during the parse
        if (fast) {
            // ---- C (pair search): Find(p) = the longest common prefix with an EARLIER position of p's bucket, the earliest one on
            // ties, capped at max_len(p) -- as a maximum over pairs it needs no order inside the buckets: every pair is compared
            // once and credited to its later position, key (length << 12 | 4095 - earlier position), by a shared-memory atomicMax.
            {   // unordered scatter into the buckets (cursors: cnt[h] ends up at the END of bucket h); lenoff is cleared meanwhile
                uint32_t* cntw = reinterpret_cast<uint32_t*>(sm.cnt);
#pragma unroll
                for (int m = 0; m < LZNT1_CHUNK / LZE_THREADS; ++m) {
                    const uint32_t h = (hh[m >> 1] >> (16 * (m & 1))) & 0xFFFFu;
                    if (h != 0xFFFFu) {
                        const uint32_t old = atomicAdd(&cntw[h >> 1], 1u << (16 * (h & 1)));
                        sm.sorted[(old >> (16 * (h & 1))) & 0xFFFFu] = (uint16_t)(tid + LZE_THREADS * m);
                    }
                }
            }
            __syncthreads();                                             // (the hashes in lenoff's first half are dead only now)
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(sm.lenoff)[tid + LZE_THREADS * i] = make_uint4(0, 0, 0, 0);
            __syncthreads();
            {
                const uint32_t* items = reinterpret_cast<const uint32_t*>(sm.ord);
                // in-tile items, one list per tile size (a warp straddling two lists runs both bodies once)
                for (uint32_t it = tid; it < basex; it += LZE_THREADS) {
                    const int ia = (int)items[it];
                    LzeEnt e0, e1;
                    lze_load(e0, sm, ia, n);
                    lze_load(e1, sm, ia + 1, n);
                    lze_pair(e0, e1, sm);
                    if (it >= base3) {
                        LzeEnt e2;
                        lze_load(e2, sm, ia + 2, n);
                        lze_pair(e0, e2, sm); lze_pair(e1, e2, sm);
                        if (it >= base4) {
                            LzeEnt e3;
                            lze_load(e3, sm, ia + 3, n);
                            lze_pair(e0, e3, sm); lze_pair(e1, e3, sm); lze_pair(e2, e3, sm);
                            if (e3.best) atomicMax(&sm.lenoff[e3.p], e3.best);
                        }
                        if (e2.best) atomicMax(&sm.lenoff[e2.p], e2.best);
                    }
                    if (e0.best) atomicMax(&sm.lenoff[e0.p], e0.best);
                    if (e1.best) atomicMax(&sm.lenoff[e1.p], e1.best);
                }
                // cross items: tile A (1..4 positions) in registers, the 4 positions of tile B streamed past it
                for (uint32_t it = basex + tid; it < n_items; it += LZE_THREADS) {
                    const uint32_t item = items[it];
                    const int ia = (int)(item & 0xFFFu), sa = (int)((item >> 12) & 3u) + 1, ib = (int)((item >> 14) & 0xFFFu);
                    LzeEnt e0, e1, e2, e3;
                    lze_load(e0, sm, ia, n);
                    lze_load(e1, sm, ia + (sa > 1 ? 1 : 0), n);
                    lze_load(e2, sm, ia + (sa > 2 ? 2 : 0), n);
                    lze_load(e3, sm, ia + (sa > 3 ? 3 : 0), n);
                    for (int u = 0; u < 4; ++u) {
                        LzeEnt f;
                        lze_load(f, sm, ib + u, n);
                        lze_pair(e0, f, sm);
                        if (sa > 1) lze_pair(e1, f, sm);
                        if (sa > 2) lze_pair(e2, f, sm);
                        if (sa > 3) lze_pair(e3, f, sm);
                        if (f.best) atomicMax(&sm.lenoff[f.p], f.best);
                    }
                    if (e0.best) atomicMax(&sm.lenoff[e0.p], e0.best);
                    if (sa > 1 && e1.best) atomicMax(&sm.lenoff[e1.p], e1.best);
                    if (sa > 2 && e2.best) atomicMax(&sm.lenoff[e2.p], e2.best);
                    if (sa > 3 && e3.best) atomicMax(&sm.lenoff[e3.p], e3.best);
                }
            }
            __syncthreads();
            // keys -> len | dist << 16 (what the parse and the emitter read), and the match bitmap
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) {
                const uint32_t v = sm.lenoff[p];
                if (v) sm.lenoff[p] = (v >> 12) | ((uint32_t)(p - (4095 - (int)(v & 0xFFFu))) << 16);
                const unsigned mm = __ballot_sync(FULL, v != 0);
                if (lane == 0) sm.M[p >> 5] = mm;
            }
            if (tid == 0) sm.M[128] = 0;
            __syncthreads();
        } else {
            // ---- ordered buckets (heavy chunks: the lazy search needs the candidates in position order) ----------------------
            // a warp's lanes hold 32 consecutive positions = one placement round: lanes with equal hashes find each other
            // here, in parallel (leader lane | rank inside the group << 5 | group size - 1 << 10, parked in ord[])
#pragma unroll 4
            for (int m = 0; m < LZNT1_CHUNK / LZE_THREADS; ++m) {
                const int p = tid + LZE_THREADS * m;
                const uint32_t h = (p < npos) ? (uint32_t)hashv[p] : 0xFFFFu;
                const unsigned peers = __match_any_sync(FULL, h);
                sm.ord[p] = (uint16_t)((uint32_t)(__ffs((int)peers) - 1) | ((uint32_t)__popc(peers & ((1u << lane) - 1)) << 5) | ((uint32_t)(__popc(peers) - 1) << 10));
            }
            __syncthreads();
            // Placement in position order by one warp, a round of 32 consecutive positions at a time: the first lane of every
            // group of equal hashes advances the bucket cursor by the size of the group with one shared-memory atomic and hands
            // the old value to the others.  Rounds are ordered (__syncwarp between their cursor updates), lanes are ordered
            // inside a round, so buckets come out ascending by position -- which the early exit and the tie rule of the search
            // rely on.  Four rounds are in flight at a time (nothing but the cursors links them).  The other warps wait at the
            // barrier; other blocks of the SM fill the issue slots meanwhile.
            if (warp == 0) {
                uint32_t* cntw = reinterpret_cast<uint32_t*>(sm.cnt);
                const int rounds = (npos + 31) >> 5;
                for (int k = 0; k < rounds; k += 4) {
                    uint32_t h[4], inf[4], got[4];
    #pragma unroll
                    for (int u = 0; u < 4; ++u) { const int p = min(32 * (k + u) + lane, LZNT1_CHUNK - 1); h[u] = hashv[p]; inf[u] = sm.ord[p]; }
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        got[u] = 0;
                        if (32 * (k + u) + lane < npos && (inf[u] & 0x3E0u) == 0)
                            got[u] = atomicAdd(&cntw[h[u] >> 1], ((inf[u] >> 10) + 1u) << (16 * (h[u] & 1)));
                        __syncwarp();                                    // cursor updates of round u before those of round u+1
                    }
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t base = (__shfl_sync(FULL, got[u], (int)(inf[u] & 31u)) >> (16 * (h[u] & 1))) & 0xFFFFu;
                        if (32 * (k + u) + lane < npos) sm.sorted[base + ((inf[u] >> 5) & 31u)] = (uint16_t)(32 * (k + u) + lane);
                    }
                }
            }
            __syncthreads();
            // cnt[h] = end offset of bucket h; begin = cnt[h-1] (0 for h == 0)

            // ---- C: match search for every position ------------------------------------------------------
            // The k-th entry of a bucket has k candidates (the earlier entries).  Threads take the entries grouped by that
            // rank, so the lanes of a warp run the candidate loop the same number of times.
            {
                uint32_t rr[4] = {0, 0, 0, 0};                           // ranks of my 16 entries, one byte each
                heavy = false;
    #pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int idx = tid + LZE_THREADS * m;
                    if (idx < npos) {
                        const uint32_t h = hashv[sm.sorted[idx]];
                        int r = idx - (int)(h ? sm.cnt[h - 1] : 0);
                        r = r < 255 ? r : 255;
                        rr[m >> 2] |= (uint32_t)r << (8 * (m & 3));
                        atomicAdd(&sm.rankcnt[r], 1u);
                    }
                }
                __syncthreads();
                {   // exclusive scan of the 256 rank counts (one per thread); sum of rank * count = candidate evaluations
                    const uint32_t v = sm.rankcnt[tid];
                    const int incl = warp_incl_scan((int)v);
                    const uint32_t wk = __reduce_add_sync(FULL, v * (uint32_t)tid);
                    if (lane == 31) { sm.scan_tmp[warp] = (uint32_t)incl; sm.work_tmp[warp] = wk; }
                    __syncthreads();
                    uint32_t woff = 0, work = 0;
                    for (int w = 0; w < warp; ++w) woff += sm.scan_tmp[w];
                    for (int w = 0; w < LZE_THREADS / 32; ++w) work += sm.work_tmp[w];
                    heavy = work > heavy_work;
                    sm.rankcnt[tid] = woff + (uint32_t)incl - v;
                }
                __syncthreads();
    #pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int idx = tid + LZE_THREADS * m;
                    if (idx < npos) sm.ord[atomicAdd(&sm.rankcnt[(rr[m >> 2] >> (8 * (m & 3))) & 0xFFu], 1u)] = (uint16_t)idx;
                }
            }
            __syncthreads();                                             // hashes are dead from here: lenoff takes their place
    #pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(sm.lenoff)[tid + LZE_THREADS * i] = make_uint4(0, 0, 0, 0);
            const int rank0_end = (int)sm.rankcnt[0];                    // entries of rank 0 have no candidate: no match
            const int defer_start = heavy ? (int)sm.rankcnt[LZE_LAZY_RANK - 1] : npos;   // rankcnt[r] = end of rank r by now
            __syncthreads();
            for (int i = defer_start + tid; i < npos; i += LZE_THREADS) { const int p = sm.sorted[sm.ord[i]]; atomicOr(&sm.U[p >> 5], 1u << (p & 31)); }
            for (int i = rank0_end + tid; i < defer_start; i += LZE_THREADS) {
                const int idx = sm.ord[i];
                const int p = sm.sorted[idx];
                int S, lim;
                lznt1_shift_limit(p, S, lim);
                const int max_len = min(n - p, (1 << S) + 2);           // mask3 (src/lznt1_compress.cpp:51,:66)
                const uint32_t key = lz_load4(sm.in, p), key2 = lz_load4(sm.in, p + 4);
                const uint32_t h = lz_hash3(key);
                int k = (h == 0) ? 0 : (int)sm.cnt[h - 1];
                int best = 0, bj = 0;
                int j = sm.sorted[k];
                uint32_t dj = lz_load4(sm.in, j);
                for (;;) {                                              // earlier entries of my bucket = earlier positions
                    const int kn = k + 1;
                    int jn = 0;
                    uint32_t djn = 0;
                    if (kn < idx) { jn = sm.sorted[kn]; djn = lz_load4(sm.in, jn); }   // next candidate: loads in flight during the compare
                    // common prefix of p and j, 4 bytes at a time; a hash collision simply ends with i < 3
                    int i2 = 0;
                    uint32_t d = dj ^ key;
                    if (!d) {
                        i2 = 4;
                        d = lz_load4(sm.in, j + 4) ^ key2;
                        while (!d) {
                            i2 += 4;
                            if (i2 >= max_len) break;
                            d = lz_load4(sm.in, p + i2) ^ lz_load4(sm.in, j + i2);
                        }
                    }
                    if (d) i2 += (__ffs((int)d) - 1) >> 3;
                    if (i2 > max_len) i2 = max_len;
                    if (i2 > best) { best = i2; bj = j; if (best == max_len) break; }
                    if (kn >= idx) break;
                    k = kn; j = jn; dj = djn;
                }
                if (best >= 3) sm.lenoff[p] = (uint32_t)best | ((uint32_t)(p - bj) << 16);
            }
            __syncthreads();
            for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) {
                const unsigned mm = __ballot_sync(FULL, sm.lenoff[p] != 0);
                if (lane == 0) sm.M[p >> 5] = mm;
            }
            if (tid == 0) sm.M[128] = 0;
            __syncthreads();
        }

        // ---- D: greedy parse p += max(1, len[p]) ------------------------------------------
        // For a match position q the parse continues at nxt(q) = first match position >= q + len[q] (everything in
        // between is literals); nxt is computed for all match positions in parallel, the visited ones are then the
        // chain from the first match -- one shared-memory load per visited match.
        {
            if (!heavy) {
                if (warp == 0) {   // nmw[w] = first word >= w of M that is not zero (128 = none)
                    int nz[4], carry = 128;
#pragma unroll
                    for (int i = 3; i >= 0; --i) { const int w = lane * 4 + i; carry = sm.M[w] ? w : carry; nz[i] = carry; }
                    int v = nz[0];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_down_sync(FULL, v, d); if (lane + d < 32) v = min(v, t); }
                    int after = __shfl_down_sync(FULL, v, 1);
                    if (lane == 31) after = 128;
#pragma unroll
                    for (int i = 0; i < 4; ++i) sm.nmw[lane * 4 + i] = (uint16_t)(nz[i] == 128 ? after : nz[i]);
                    if (lane == 0) sm.nmw[128] = 128;
                }
                __syncthreads();
                uint16_t* cur = sm.sorted;                                                        // dead after stage C
                uint16_t* nxt2 = sm.ord;                                                          // dead after stage C as well
                auto first_match_from = [&](int e) -> int {
                    if (e >= LZNT1_CHUNK) return LZNT1_CHUNK;
                    int w = e >> 5;
                    uint32_t bits = sm.M[w] & (0xffffffffu << (e & 31));
                    if (!bits) { w = sm.nmw[w + 1]; bits = sm.M[w]; }                              // M[128] == 0
                    return bits ? (w * 32 + __ffs((int)bits) - 1) : LZNT1_CHUNK;
                };
                for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS)
                    cur[p] = (uint16_t)(((sm.M[p >> 5] >> (p & 31)) & 1) ? first_match_from(p + (int)(sm.lenoff[p] & 0xFFFF)) : LZNT1_CHUNK);
                __syncthreads();
                // two rounds of pointer doubling shorten the serial part of the chase by 4: nxt2 = nxt o nxt, nxt4 = nxt2 o nxt2
                for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) { const int t = cur[p]; nxt2[p] = (uint16_t)(t < LZNT1_CHUNK ? cur[t] : LZNT1_CHUNK); }
                __syncthreads();
                for (int p = tid; p < LZNT1_CHUNK; p += LZE_THREADS) { const int t = nxt2[p]; sm.nxt4[p] = (uint16_t)(t < LZNT1_CHUNK ? nxt2[t] : LZNT1_CHUNK); }
                __syncthreads();
                if (tid == 0) {     // every 4th visited match: one shared-memory load per link
                    int q = first_match_from(0), k = 0;
                    while (q < LZNT1_CHUNK) { sm.chain[k++] = (uint16_t)q; q = sm.nxt4[q]; }
                    sm.nchain = (uint32_t)k;
                }
                __syncthreads();
                for (uint32_t k = tid; k < sm.nchain; k += LZE_THREADS) {                         // the three in between, in parallel
                    const int q = sm.chain[k], q1 = cur[q], q2 = nxt2[q];
                    lze_cover(sm.C, q, (int)(sm.lenoff[q] & 0xFFFF));
                    if (q1 < LZNT1_CHUNK) lze_cover(sm.C, q1, (int)(sm.lenoff[q1] & 0xFFFF));
                    if (q2 < LZNT1_CHUNK) {
                        lze_cover(sm.C, q2, (int)(sm.lenoff[q2] & 0xFFFF));
                        const int q3 = cur[q2];
                        if (q3 < LZNT1_CHUNK) lze_cover(sm.C, q3, (int)(sm.lenoff[q3] & 0xFFFF));
                    }
                }
            } else if (warp == 0) {
                // Heavy chunk: one warp walks the parse; a deferred position is searched when (and only if) it is reached.
                int p = 0;
                while (p < n) {
                    // first position >= p that has a match or a deferred search (everything before it is a literal)
                    int q = LZNT1_CHUNK;
                    for (int w0 = p >> 5; w0 < 128; w0 += 32) {
                        const int w = w0 + lane;
                        uint32_t bits = (w < 128) ? (sm.M[w] | sm.U[w]) : 0u;
                        if (w == (p >> 5)) bits &= 0xffffffffu << (p & 31);
                        const unsigned nz = __ballot_sync(FULL, bits != 0);
                        if (nz) {
                            const int f = __ffs((int)nz) - 1;
                            q = (w0 + f) * 32 + __ffs((int)__shfl_sync(FULL, bits, f)) - 1;
                            break;
                        }
                    }
                    if (q >= LZNT1_CHUNK) break;
                    uint32_t lo = sm.lenoff[q];
                    if ((sm.U[q >> 5] >> (q & 31)) & 1) {
                        // Find(q) by the warp: candidates = entries of q's bucket before q, 32 per round, ascending
                        int S, lim;
                        lznt1_shift_limit(q, S, lim);
                        const int max_len = min(n - q, (1 << S) + 2);
                        const uint32_t key = lz_load4(sm.in, q);
                        const uint32_t h = lz_hash3(key);
                        const int kend = (int)sm.cnt[h];
                        uint32_t bestkey = 0;                               // len << 12 | (4095 - j): longest, then earliest
                        for (int kb = (h == 0) ? 0 : (int)sm.cnt[h - 1]; kb < kend; kb += 32) {
                            const int k = kb + lane;
                            const int j = (k < kend) ? (int)sm.sorted[k] : LZNT1_CHUNK;
                            const bool valid = j < q;
                            int i2 = 0;
                            if (valid) {
                                uint32_t d = lz_load4(sm.in, j) ^ key;
                                while (!d) {
                                    i2 += 4;
                                    if (i2 >= max_len) break;
                                    d = lz_load4(sm.in, q + i2) ^ lz_load4(sm.in, j + i2);
                                }
                                if (d) i2 += (__ffs((int)d) - 1) >> 3;
                                if (i2 > max_len) i2 = max_len;
                                if (i2 >= 3) bestkey = max(bestkey, ((uint32_t)i2 << 12) | (uint32_t)(4095 - j));
                            }
                            // ascending candidates: the first round that reaches max_len holds the earliest such candidate
                            if (__any_sync(FULL, !valid || i2 == max_len)) break;
                        }
                        bestkey = __reduce_max_sync(FULL, bestkey);
                        lo = bestkey ? ((bestkey >> 12) | ((uint32_t)(q - (4095 - (int)(bestkey & 4095u))) << 16)) : 0u;
                        __syncwarp();
                        if (lane == 0) {
                            sm.lenoff[q] = lo;
                            sm.U[q >> 5] &= ~(1u << (q & 31));
                            if (lo) sm.M[q >> 5] |= 1u << (q & 31);
                        }
                        __syncwarp();
                    }
                    if (lo) { if (lane == 0) lze_cover(sm.C, q, (int)(lo & 0xFFFF)); p = q + (int)(lo & 0xFFFF); }
                    else p = q + 1;
                }
            }
            __syncthreads();
            // visited bitmap: every position of the chunk that is not inside a taken match starts an item
            if (tid < 128) {
                const int lo = 32 * tid;
                const uint32_t inside = n >= lo + 32 ? 0xffffffffu : (n > lo ? (0xffffffffu >> (32 - (n - lo))) : 0u);
                sm.V[tid] = ~sm.C[tid] & inside;
            }
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
