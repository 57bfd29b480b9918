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
constexpr int LZX_FIND_THREADS = 512;         // lz_find_kernel: 16 warps on one 64 KiB segment -- the fewer segments an SM has in flight, the more of their
                                               // links and bytes stay in L1 / L2 (Xpress-Huffman chains roam the whole 64 KiB)
constexpr int LZX_FIND_MINB = 3;

// the reference's progressive hash after three updates (XpressDictionary.h:58-61 with HashBits = 15): the first byte's high bits fall off
__device__ __forceinline__ uint32_t lzx_hash(uint32_t key3) { return (((key3 & 0xFFu) << 10) ^ ((key3 >> 3) & 0x1FE0u) ^ ((key3 >> 16) & 0xFFu)) & 0x7FFFu; }

// One warp per work item: `piece` positions of a unit plus `window` positions of warm-up before them.  T[h] = 1 + (position -
// warm-up start) of the latest position with hash h among those already passed, 0 = none; TE must hold piece + window.
// 32 positions per round: the nearest earlier lane with the same hash (__match_any_sync) wins, else the table; the highest lane of
// every hash group updates the table.  dist[] (one entry per position, indexed like the units' segments: seg_first[u] * LZX_SEG + p)
// receives p - prev(p), or 0 when there is none within the window.
// The rounds form one dependent chain through the table, so nothing else may wait inside a round: the input comes through a
// shared-memory stage of LZX_STAGE bytes that is refilled from registers loaded one stage ahead (aligned 16-byte loads, the
// whole warp), and every round reads its 3-byte keys from there.
constexpr uint32_t LZX_STAGE = 2048;           // positions per stage (a multiple of 512); the stage holds 16 bytes more for the keys at its end
template <typename TE>
__global__ void __launch_bounds__(32)
lz_prev_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
               const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ work_counter, uint16_t* __restrict__ dist,
               uint32_t window, uint32_t piece)
{
    MSC_DYN_SMEM(smem_raw);
    TE* T = reinterpret_cast<TE*>(smem_raw);
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw + 32768u * sizeof(TE));         // LZX_STAGE + 16 bytes (+ slack for the look-ahead)
    const int lane = (int)threadIdx.x;
    const uint32_t per_seg = LZX_SEG / piece;
    const uint32_t total = seg_first[count] * per_seg;
    constexpr int VPL = LZX_STAGE / 512;                               // 16-byte vectors per lane and stage
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
        // the unit as aligned 16-byte vectors: vector v holds bytes [16 v - skew, 16 v - skew + 16) of the unit
        const uintptr_t addr = reinterpret_cast<uintptr_t>(in + in_off[u]);
        const uint4* vbase = reinterpret_cast<const uint4*>(addr & ~static_cast<uintptr_t>(15));
        const uint32_t skew = (uint32_t)(addr & 15), vlast = (skew + n - 1u) >> 4;      // last vector holding a byte of the unit
        uint16_t* dst = dist + (uint64_t)seg_first[u] * LZX_SEG;
        for (uint32_t i = (uint32_t)lane; i < 32768u * sizeof(TE) / 16u; i += 32) reinterpret_cast<uint4*>(T)[i] = make_uint4(0, 0, 0, 0);
        // stage k covers unit bytes [b0 + k * LZX_STAGE, ... + LZX_STAGE + 16), b0 = the 16-byte granule of w0
        const uint32_t v0 = (skew + w0) >> 4;                          // first vector of stage 0
        uint4 nx[VPL + 1];
        auto fetch = [&](uint32_t vfirst) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) { const uint32_t v = vfirst + (uint32_t)(j * 32 + lane); nx[j] = v <= vlast ? __ldg(vbase + v) : make_uint4(0, 0, 0, 0); }
            const uint32_t v = vfirst + (uint32_t)(VPL * 32);
            nx[VPL] = (lane == 0 && v <= vlast) ? __ldg(vbase + v) : make_uint4(0, 0, 0, 0);
        };
        fetch(v0);
        __syncwarp();
        for (uint32_t sb = 0;; ++sb) {                                 // stages
            const uint32_t vfirst = v0 + sb * (LZX_STAGE / 16);
            const uint32_t pb = vfirst * 16u - skew;                   // unit position of stage byte 0 (may be "negative" for the first stage: wraps, unused)
#pragma unroll
            for (int j = 0; j < VPL; ++j) reinterpret_cast<uint4*>(stage)[j * 32 + lane] = nx[j];
            if (lane == 0) reinterpret_cast<uint4*>(stage)[VPL * 32] = nx[VPL];
            __syncwarp();
            const uint32_t pend = pb + LZX_STAGE;                      // positions of this stage: [max(pb, w0), min(pend, s1))
            if (pend < s1) fetch(vfirst + LZX_STAGE / 16);              // the next stage is on its way while this one is processed
            const uint32_t first = (sb == 0) ? w0 : pb;
            // the hash of a round is computed one round ahead: it does not depend on the table, the chain through the table does not wait for it
            auto hash_at = [&](uint32_t q) -> uint32_t { const uint32_t o = q - pb; return lzx_hash(__funnelshift_r(stage[o >> 2], stage[(o >> 2) + 1], 8u * (o & 3u))); };
            uint32_t h_next = hash_at(first + (uint32_t)lane);
            for (uint32_t p0 = first; p0 < min(pend, s1); p0 += 32) {
                const uint32_t p = p0 + (uint32_t)lane;
                const bool on = p + 3 <= n && p < pend;
                const uint32_t h = h_next;
                h_next = hash_at(p + 32u);                             // (beyond the stage on its last round: read from the slack, never used)
                const uint32_t mine = p - w0 + 1u;
                // Fast path, no two lanes with the same hash: everybody reads the table, writes itself, and reads itself back.
                // (__match_any_sync costs one pass per distinct value -- 32 of them on ordinary data -- so it is kept for the
                // rounds where some lane lost the write.)
                const uint32_t told = on ? (uint32_t)T[h] : 0u;
                __syncwarp();
                if (on) T[h] = (TE)mine;
                __syncwarp();
                const bool lost = on && (uint32_t)T[h] != mine;
                uint32_t d = (on && told) ? mine - told : 0u;
                if (__any_sync(FULL, lost)) {
                    const unsigned peers = __match_any_sync(FULL, on ? h : (0x8000u | (uint32_t)lane));
                    const unsigned below = peers & ((1u << lane) - 1u);
                    if (on && below) d = (uint32_t)(lane - (31 - __clz((int)below)));
                    __syncwarp();
                    if (on && (peers >> lane) == 1u) T[h] = (TE)mine;       // the highest lane of every hash group is the latest position
                    __syncwarp();
                }
                if (d > window) d = 0;
                if (p >= s0 && p < s1 && p < pend) dst[p] = (uint16_t)d;
            }
            __syncwarp();
            if (pend >= s1) break;
        }
    }
}

// What Find needs to know about a unit.
struct LzxUnit {
    LzfSrc S;                       // the unit's bytes as aligned words
    const uint8_t* src;
    const uint16_t* dd;             // its links (lz_prev_kernel)
    uint32_t n, window, chunk;      // chunk = 0: a match may run to the end of the unit (plain Xpress); otherwise the caller clips it at
                                    // the end of the chunk it starts in and does not ask at all with fewer than 3 bytes of chunk left
                                    // (Xpress-Huffman: src/xpress_huff_compress.cpp:90-93)
    __device__ __forceinline__ LzxUnit(const uint8_t* s, uint32_t n_, const uint16_t* d, uint32_t w, uint32_t c) : S(s, n_), src(s), dd(d), n(n_), window(w), chunk(c) {}
};
constexpr uint32_t LZX_UNKNOWN = 0xFFFFFFFFu;  // lenoff[] entry of a position nobody has asked about

// find(p), one thread: (len - 1) << 16 | offset, or 0 when Find would return less than 3 (or is not asked).  A candidate can only
// beat the current length `len` if it agrees on byte `len`: one probe word per link decides (the three bytes of the key while
// len is still 2, afterwards the four bytes ending at index len); a full compare only follows a hit.
__device__ __forceinline__ uint32_t lzx_find_one(const LzxUnit& U, uint32_t p)
{
    if (p + 3 > U.n) return 0;
    const uint32_t rem = U.chunk ? U.chunk - (p % U.chunk) : 0xFFFFFFFFu;    // Find itself does not know about chunks ...
    const uint32_t maxlen = min(U.n - p - 1u, LZX_CAP);                      // GetMatchLength never reaches the last byte (:74-93)
    if (maxlen < 3u || rem < 3u) return 0;
    uint32_t len = 2, off = 0, x = p;
    uint32_t po = 0, mask = 0xFFFFFFu, pk = lzf_ld4(U.S, p) & mask;
    for (uint32_t c = 0; c < LZX_CHAIN; ++c) {
        const uint32_t d = U.dd[x];
        if (!d || p - x + d > U.window) break;
        x -= d;
        if ((lzf_ld4(U.S, x + po) & mask) != pk) continue;                   // hash collisions and shorter matches only use up a link
        const uint32_t l = lzf_match_len(U.S, U.src, p, x, maxlen);
        if (l > len) {
            len = l; off = p - x;
            if (len >= LZX_NICE || len == maxlen) break;                     // (nothing is longer than maxlen: the walk could only waste links)
            po = len - 3u; mask = 0xFFFFFFFFu; pk = lzf_ld4(U.S, p + po);
        }
    }
    len = min(len, rem);                                                     // ... the caller clips what it found (:93)
    return len >= 3u ? ((len - 1u) << 16) | off : 0u;
}

// Find only where a greedy parse goes.  The input is cut into sub-tiles of LZX_SUB positions; a lane owns a sub-tile and runs the
// greedy parse p += max(1, len(find(p))) from the sub-tile's FIRST position -- an assumption: the true parse enters wherever the
// match before ends -- asking Find only at the positions it steps on (a quarter of all positions on text).  Greedy parses that
// meet stay together, so the true parse, entering elsewhere, follows steps of its own only until it stands on a token start of
// the assumed one; those few extra questions are asked when the sub-tiles are stitched (lz_stitch_kernel, lzx_merge_parse).
// Out: lenoff[] (indexed like dist[]; pre-set to LZX_UNKNOWN) at the visited positions; per sub-tile the bitmap of token starts
// (two words, laid out so that the LZX_SUB / 32 words of consecutive sub-tiles form the per-tile bitmap) and the unit position
// where the assumed parse leaves the sub-tile.
// Persistent blocks pull 64 KiB segments; the 32 lanes of a warp own 32 consecutive sub-tiles (= one tile).  A lane is in one of
// three states: ST_HOP (follow a link, probe), ST_CMP (the probe hit: measure the candidate), ST_NEW (a position is answered:
// store, step, start the next).  Every round the WARP runs the code of one state only -- the one most of its lanes are in -- so the
// rare, long pieces (compare, next position) run for many lanes at once instead of once per link round with a handful of lanes.
constexpr uint32_t LZX_SUB = 64;
template <int MINB>
__global__ void __launch_bounds__(LZX_FIND_THREADS, MINB)
lz_find_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
               const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ work_counter, const uint16_t* __restrict__ dist,
               uint32_t* __restrict__ lenoff, uint32_t* __restrict__ visited_sub, uint32_t* __restrict__ exit_sub, uint32_t window, uint32_t chunk)
{
    __shared__ uint32_t s_work, s_next;
    const int tid = (int)threadIdx.x;
    const uint32_t total = seg_first[count];
    for (;;) {
        if (tid == 0) s_work = atomicAdd(work_counter, 1u);
        __syncthreads();                                               // (everybody is done with the segment before)
        const uint32_t w = s_work;
        if (tid == 0) s_next = LZX_FIND_THREADS;                       // sub-tiles beyond the first round are handed out on demand
        __syncthreads();
        if (w >= total) break;
        uint32_t lo = 0, hi = count;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seg_first[mid] <= w) lo = mid; else hi = mid; }
        const uint32_t u = lo, n = in_len[u];
        const uint32_t s0 = (w - seg_first[u]) * LZX_SEG, s1 = min(s0 + LZX_SEG, n);
        const uint8_t* src = in + in_off[u];
        const LzfSrc S(src, n);
        const uint64_t ubase = (uint64_t)seg_first[u] * LZX_SEG;
        const uint16_t* dd = dist + ubase;
        uint32_t* out = lenoff + ubase;

        enum { ST_HOP = 0, ST_CMP = 1, ST_NEW = 2, ST_END = 3 };
        uint32_t b = s0 + (uint32_t)tid * LZX_SUB;                     // the sub-tile in hand: [b, e)
        uint32_t e = min(b + LZX_SUB, s1);
        uint32_t p = b, vlo = 0, vhi = 0;
        uint32_t x = 0, len = 0, off = 0, po = 0, mask = 0, pk = 0, links = 0, maxlen = 0, rem = 0;
        uint32_t st = ST_NEW, res = 0;
        bool fresh = true;                                             // ST_NEW without an answer to store (a sub-tile's first position)
        if (b >= s1) st = ST_END;
        for (;;) {
            // the census of the three states in one warp reduction: 8 bits per state (ST_END contributes nothing)
            const uint32_t census = __reduce_add_sync(FULL, st < ST_END ? 1u << (8u * st) : 0u);
            if (!census) break;
            const int c_hop = (int)(census & 0xFFu), c_cmp = (int)((census >> 8) & 0xFFu), c_new = (int)(census >> 16);
            if (c_hop >= c_cmp && c_hop >= c_new) {
#pragma unroll
                for (int rep = 0; rep < 2; ++rep)                      // two links per round: half the votes per link
                if (st == ST_HOP) {
                    const uint32_t d = dd[x];
                    if (!d || p - x + d > window || links == 0u) { res = len >= 3u ? ((min(len, rem) - 1u) << 16) | off : 0u; st = ST_NEW; }
                    else {
                        x -= d; --links;
                        if ((lzf_ld4(S, x + po) & mask) == pk) st = ST_CMP;        // (hash collisions and shorter matches only use up a link)
                    }
                }
            } else if (c_cmp >= c_new) {
                if (st == ST_CMP) {
                    const uint32_t l = lzf_match_len(S, src, p, x, maxlen);
                    st = ST_HOP;
                    if (l > len) {
                        len = l; off = p - x;
                        if (len >= LZX_NICE || len == maxlen) {        // (nothing is longer than maxlen: the walk could only waste links)
                            res = ((min(len, rem) - 1u) << 16) | off;  // the caller clips what Find found (:93); rem >= 3
                            st = ST_NEW;
                        } else { po = len - 3u; mask = 0xFFFFFFFFu; pk = lzf_ld4(S, p + po); }
                    }
                }
            } else {
                if (st == ST_NEW) {
                    for (;;) {
                        if (!fresh) {                                  // answer in hand: store it, mark the token start, step
                            out[p] = res;
                            const uint32_t k = p - b;
                            if (k < 32u) vlo |= 1u << k; else vhi |= 1u << (k - 32u);
                            p += res ? (res >> 16) + 1u : 1u;
                        }
                        fresh = false;
                        if (p >= e) {                                  // the assumed parse leaves the sub-tile
                            const uint64_t sub = (ubase + b) / LZX_SUB;
                            visited_sub[2 * sub] = vlo; visited_sub[2 * sub + 1] = vhi; exit_sub[sub] = p;
                            b = s0 + atomicAdd(&s_next, 1u) * LZX_SUB;     // whoever is done first takes the next one
                            if (b >= s1) { st = ST_END; break; }
                            e = min(b + LZX_SUB, s1); p = b; vlo = vhi = 0;
                        }
                        bool ask = p + 3u <= n;
                        if (ask) {
                            rem = chunk ? chunk - (p % chunk) : 0xFFFFFFFFu;          // Find itself does not know about chunks ...
                            maxlen = min(n - p - 1u, LZX_CAP);                         // GetMatchLength never reaches the last byte (:74-93)
                            ask = maxlen >= 3u && rem >= 3u;
                        }
                        if (ask) { x = p; len = 2; off = 0; po = 0; mask = 0xFFFFFFu; pk = lzf_ld4(S, p) & mask; links = LZX_CHAIN; st = ST_HOP; fresh = false; break; }
                        res = 0;                                       // Find is not even asked here: a literal on the spot
                    }
                    if (st == ST_HOP) fresh = false;
                }
            }
        }
    }
}

// Stitches the sub-tile parses of lz_find_kernel into the greedy parse of every tile of LZF_TILE positions AS IF a token started at
// the tile's first position: one thread per tile walks its sub-tiles in order, entering each where the one before was left;
// inside a sub-tile it steps (asking Find for positions nobody asked about, and recording the answer) until it stands on a token
// start of the sub-tile's own parse, whose remainder and exit it then takes over.  Out, per tile (global tile number
// seg_first[u] * (LZX_SEG / LZF_TILE) + position / LZF_TILE): the token-start bitmap visited0[LZF_TILE / 32] (in place: the
// sub-tile words ARE the tile's words) and exit0 = the unit position where that parse leaves the tile.
__global__ void __launch_bounds__(128)
lz_stitch_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                 const uint32_t* __restrict__ seg_first, const uint16_t* __restrict__ dist, uint32_t* __restrict__ lenoff,
                 uint32_t* __restrict__ visited0, const uint32_t* __restrict__ exit_sub, uint32_t* __restrict__ exit0, uint32_t window, uint32_t chunk)
{
    constexpr uint32_t TPS = LZX_SEG / LZF_TILE, SPT = LZF_TILE / LZX_SUB;       // tiles per segment, sub-tiles per tile
    const uint64_t ntiles = (uint64_t)seg_first[count] * TPS;
    for (uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gt < ntiles; gt += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t sg = (uint32_t)(gt / TPS);
        uint32_t lo = 0, hi = count;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seg_first[mid] <= sg) lo = mid; else hi = mid; }
        const uint32_t u = lo, n = in_len[u];
        const uint64_t ubase = (uint64_t)seg_first[u] * LZX_SEG;
        const uint32_t t0 = (uint32_t)(gt * LZF_TILE - ubase);
        if (t0 >= n) { exit0[gt] = t0; continue; }
        const uint32_t tend = min(t0 + LZF_TILE, n);
        const LzxUnit U(in + in_off[u], n, dist + ubase, window, chunk);
        uint32_t* lo_u = lenoff + ubase;
        uint32_t e = t0;                                               // where the parse enters the next sub-tile
        for (uint32_t k = 0; k < SPT; ++k) {
            const uint32_t sb = t0 + k * LZX_SUB;
            uint32_t* vw = visited0 + (gt * SPT + k) * 2;
            if (sb >= tend) { vw[0] = 0; vw[1] = 0; continue; }
            const uint32_t se = min(sb + LZX_SUB, tend);
            if (e >= se) { vw[0] = 0; vw[1] = 0; continue; }           // a match jumps over the sub-tile
            const unsigned long long v = (unsigned long long)vw[0] | ((unsigned long long)vw[1] << 32);
            unsigned long long mine = 0;
            uint32_t p = e;
            bool met = false;
            while (p < se) {
                const uint32_t kbit = p - sb;
                if ((v >> kbit) & 1ull) { met = true; mine |= v & (~0ull << kbit); break; }
                uint32_t r = lo_u[p];
                if (r == LZX_UNKNOWN) { r = lzx_find_one(U, p); lo_u[p] = r; }
                mine |= 1ull << kbit;
                p += r ? (r >> 16) + 1u : 1u;
            }
            e = met ? exit_sub[(ubase + sb) / LZX_SUB] : p;
            vw[0] = (uint32_t)mine; vw[1] = (uint32_t)(mine >> 32);
        }
        exit0[gt] = e;
    }
}

// The true parse of a tile from the parse that assumed a token at the tile's first position (lz_stitch_kernel).  In: sm.lenoff = the
// tile's lenoff[] entries (LZX_UNKNOWN where nobody asked; the forced literals already zeroed), sm.V = visited0 of the tile,
// sm.next_p = where the parse enters (same coordinates as t0), exit0 of the tile in those coordinates; fe = position below which
// every position is a literal; U + pos0 = the unit and the unit position of the tile's first position, for the questions nobody has
// asked yet.  One thread steps from the entry until it stands on a token start of the assumed parse that lies at or beyond fe
// (from there on the two agree) or leaves the tile.  Out: sm.V, sm.prefV, sm.next_p as lzf_parse_tile leaves them; sm.lenoff
// holds an answer at every token start.  Must be called by all LZF_THREADS threads.
__device__ __forceinline__ void lzx_merge_parse(LzFrontSmem& sm, uint32_t n, uint32_t t0, uint32_t exit0, uint32_t fe, const LzxUnit& U, uint32_t pos0)
{
    const int tid = (int)threadIdx.x, lane = tid & 31;
    const uint32_t tend_r = min(t0 + LZF_TILE, n) - t0;
    if (tid < 32) {
        const uint32_t start_r = sm.next_p > t0 ? sm.next_p - t0 : 0u;
        const uint32_t fe_r = fe > t0 ? fe - t0 : 0u;
        // lane 0 walks; the words it passes are rebuilt from its steps, the words beyond the meeting point keep the assumed parse
        uint32_t p = start_r, m = LZF_TILE;                           // m = meeting point (LZF_TILE: none inside the tile)
        uint32_t w_cur = 0xFFFFFFFFu, bits = 0;
        if (lane == 0 && start_r < tend_r) {
            w_cur = start_r >> 5;
            while (p < tend_r) {
                if (p >= fe_r && ((sm.V[p >> 5] >> (p & 31)) & 1)) { m = p; break; }
                if ((p >> 5) != w_cur) { sm.K[w_cur] = bits; bits = 0; for (uint32_t w2 = w_cur + 1; w2 < (p >> 5); ++w2) sm.K[w2] = 0; w_cur = p >> 5; }
                bits |= 1u << (p & 31);
                uint32_t lo = sm.lenoff[p];
                if (lo == LZX_UNKNOWN) { lo = lzx_find_one(U, pos0 + p); sm.lenoff[p] = lo; }
                p += lo ? (lo >> 16) + 1u : 1u;
            }
            sm.K[w_cur] = bits;
        }
        m = __shfl_sync(FULL, m, 0);
        p = __shfl_sync(FULL, p, 0);
        const uint32_t wc = __shfl_sync(FULL, w_cur, 0);              // last word the walk wrote; later words were not stepped on
        __syncwarp();
        const uint32_t w_first = start_r >> 5, w_meet = m >> 5;       // m == LZF_TILE: w_meet = LZF_TILE / 32, beyond the last word
#pragma unroll
        for (int k = 0; k < LZF_TILE / 1024; ++k) {
            const uint32_t w = (uint32_t)(k * 32 + lane);
            const uint32_t kv = (start_r < tend_r && w >= w_first && w <= wc) ? sm.K[w] : 0u;
            uint32_t v;
            if (start_r >= tend_r) v = 0;                              // a long match skips the tile
            else if (w < w_meet) v = kv;
            else if (w == w_meet) v = kv | (sm.V[w] & (0xFFFFFFFFu << (m & 31)));
            else v = sm.V[w];
            sm.V[w] = v;
        }
        if (lane == 0 && start_r < tend_r) sm.next_p = (m < (uint32_t)LZF_TILE) ? exit0 : t0 + max(p, tend_r);
        __syncwarp();
        const uint32_t a = (uint32_t)__popc(sm.V[2 * lane]), b = (uint32_t)__popc(sm.V[2 * lane + 1]);
        const uint32_t incl = (uint32_t)warp_incl_scan((int)(a + b));
        sm.prefV[2 * lane] = incl - a - b;
        sm.prefV[2 * lane + 1] = incl - b;
    }
    __syncthreads();
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
