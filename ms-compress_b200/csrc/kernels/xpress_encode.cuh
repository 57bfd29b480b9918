// xpress_encode.cuh -- plain Xpress (LZ77) encoder, one 256-thread block per buffer.
//
// Replaces the reference's xpress_compress (src/xpress_compress.cpp:240-347) and its use of
// XpressDictionary<0x2000> (:28).  Matching / greedy parse: lz_front.cuh (window 8192).  Emission is done tile by
// tile with three running prefix sums instead of the reference's serial output cursor:
//   item bytes   -- literal 1; match 2 (+1 when it opens a new shared length nibble byte, +1 extension byte,
//                   +2 for a 16-bit length), src/xpress_compress.cpp:274-315
//   item count   -- a 4-byte flag word (MSB first, 1 = match) precedes every 32 items (:317-324), so item k sits at
//                   4 * (k / 32 + 1) + (bytes of the items before it)
//   nibble users -- the q-th match with length >= 10 writes the low (q even) or high (q odd) nibble of a byte
//                   that the even one allocates right after its own 2-byte symbol (:281-290)
// The stream ends inside a flag word whose unused bits are 1, which costs an extra FFFFFFFF word when the item
// count is a multiple of 32 (:343); empty input encodes as FF FF FF FF (:253-259).
#pragma once
#include "common.cuh"
#include "lz_front.cuh"
#include "lz_exact.cuh"

namespace msc {

constexpr int XPE_THREADS = LZF_THREADS;
constexpr int XPE_PER_THREAD = LZF_TILE / XPE_THREADS;     // 8 consecutive positions per thread

// Restart point of a plain-Xpress stream (side-band, never part of the stream): everything a decoder needs to resume at an item
// boundary.  One entry per segment of seg_tiles * LZF_TILE input bytes; entry 0 (stream start) also carries the entry count.
// Layout == mscomp_gpu_xpress_seg (include/mscomp_gpu.h).
struct XpressSeg {
    uint32_t in_pos;      // compressed offset of the next byte to read (the flag word itself when flag_used == 0)
    uint32_t out_pos;     // uncompressed position of the next item
    uint32_t flag_pos;    // compressed offset of the flag word in force (valid when flag_used != 0)
    uint32_t flag_used;   // items of that flag word already consumed (0..31)
    uint32_t half_pos;    // compressed offset of the byte whose high nibble is still unused, 0xFFFFFFFF = none
    uint32_t nseg;        // entry 0: number of entries of this stream
};

struct XpEncSmem {
    LzFrontSmem lz;
    uint32_t flagw[LZF_TILE / 32 + 2];       // flag words touched by this tile (first one may be carried in)
    uint32_t flagpos[LZF_TILE / 32 + 2];     // their byte positions in the output
    uint32_t nibpos[LZF_TILE];               // half-byte position of the even-ranked nibble users of this tile
    unsigned long long wsum[XPE_THREADS / 32];
    // carried across tiles
    uint32_t items, bytes, nibs;             // totals so far (bytes = item bytes only, flag words not included)
    uint32_t pend_half;                      // position of the half byte waiting for its high nibble (nibs odd)
    uint32_t cur_flags, cur_flagpos;         // the partially filled flag word
    int status;
    // reference-parse variant (MODE 2)
    uint32_t forced_end;                     // positions below it are literals whatever was found there (lzx_forced_literals)
    uint32_t big_pos, big_len;               // the tile's visited match of LZX_CAP or more bytes (relative position, LZF_TILE = none), its true length
    uint32_t red;
};

// bytes of a match item, given whether it has to allocate the shared half byte
__device__ __forceinline__ uint32_t xpe_match_bytes(uint32_t lm3, bool new_half)
{
    uint32_t b = 2;
    if (lm3 >= 7) { b += new_half ? 1u : 0u; if (lm3 - 7 >= 15) { b += 1; if (lm3 - 22 >= 255) b += lm3 <= 0xFFFFu ? 2u : 6u; } }
    return b;
}

// Two-pass variant for large buffers: match finding is the expensive part and needs no stream state, so it runs per 64 KiB
// segment on its own block (xpress_match_kernel: matches and parse never cross a segment, the results of every tile -- lenoff,
// visited bitmap, carried-in position -- go to scratch); the emission, which IS sequential along the stream (flag words,
// shared nibbles), then only replays tiles from scratch (PRE = true).  One buffer no longer costs one block its whole
// match-finding time: few large buffers, down to the single buffer of the one-shot call, use the whole GPU for the expensive part.
constexpr uint32_t XPE_SEG = 65536;                                  // match-finding segment (a multiple of LZF_TILE)
constexpr uint32_t XPE_SEG_TILES = XPE_SEG / LZF_TILE;
struct XpPre {                                                       // scratch of the two-pass variant, indexed by global tile number
    const uint32_t* seg_first;                                       // [count + 1] first global segment of every buffer
    uint32_t* lenoff;                                                // [tiles][LZF_TILE]
    uint32_t* visited;                                               // [tiles][LZF_TILE / 32]
    uint32_t* next_p;                                                // [tiles] position (relative to the segment) where the parse enters the tile
    const uint16_t* dist;                                            // reference parse (MODE 2): the links, for questions nobody has asked yet
};

// MODE 0: matching and emission in one kernel; 1: emission from the tiles of xpress_match_kernel; 2: the reference's parse -- emission
// from the answers of lz_find_kernel (lz_exact.cuh), the greedy parse runs here, along the whole buffer: byte-identical to the
// reference's xpress_compress.
template <int MODE>
__device__ __forceinline__ void xpress_encode_body(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                     const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                     const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                     uint32_t* __restrict__ out_len, int32_t* __restrict__ status,
                     uint32_t seg_tiles, XpressSeg* __restrict__ index, uint32_t index_stride, const XpPre pre)
{
    MSC_DYN_SMEM(smem_raw);
    XpEncSmem& sm = *reinterpret_cast<XpEncSmem*>(smem_raw);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool PRE = MODE == 1, EXACT = MODE == 2;

    for (uint32_t u = blockIdx.x; u < count; u += gridDim.x) {
        const uint8_t* src = in + in_off[u];
        const uint32_t n = in_len[u], cap = out_cap[u];
        uint8_t* o = out + out_off[u];
        if (MODE == 0) for (int i = tid; i < (1 << LZF_HBITS); i += XPE_THREADS) sm.lz.H[i] = 0;
        if (tid == 0) {
            sm.forced_end = 0; sm.big_pos = LZF_TILE; sm.big_len = 0;
            sm.lz.next_p = 0; sm.lz.ntok = 0;
            sm.items = 0; sm.bytes = 0; sm.nibs = 0; sm.pend_half = 0; sm.cur_flags = 0; sm.cur_flagpos = 0;
            sm.status = (cap < 4) ? ST_BUF : ST_OK;
        }
        __syncthreads();

        XpressSeg* idx = index ? index + (uint64_t)u * index_stride : nullptr;
        if (idx && tid == 0) {
            const uint32_t segb = seg_tiles * LZF_TILE;
            const uint32_t nseg = n ? (n + segb - 1) / segb : 1u;
            idx[0] = XpressSeg{0u, 0u, 0u, 0u, 0xFFFFFFFFu, nseg < index_stride ? nseg : index_stride};
        }
        for (uint32_t t0 = 0; t0 < n && sm.status == ST_OK; t0 += LZF_TILE) {
            uint32_t tile = 0;
            if (PRE) {                                                       // where the parse enters this tile (known from pass 1)
                tile = (pre.seg_first[u] + t0 / XPE_SEG) * XPE_SEG_TILES + (t0 % XPE_SEG) / LZF_TILE;
                if (tid == 0) sm.lz.next_p = (t0 / XPE_SEG) * XPE_SEG + pre.next_p[tile];
                __syncthreads();
            }
            if (idx && tid == 0 && t0 && (t0 / LZF_TILE) % seg_tiles == 0 && t0 / (LZF_TILE * seg_tiles) < index_stride) {
                const uint32_t k = sm.items, c = k & 31u;                    // next item; the state below is final for everything before it
                XpressSeg e;
                e.in_pos = c ? sm.bytes + 4 * ((k >> 5) + 1) : sm.cur_flagpos;
                e.out_pos = sm.lz.next_p;
                e.flag_pos = sm.cur_flagpos;
                e.flag_used = c;
                e.half_pos = (sm.nibs & 1u) ? sm.pend_half : 0xFFFFFFFFu;
                e.nseg = 0;
                idx[t0 / (LZF_TILE * seg_tiles)] = e;
            }
            if (PRE) {
                const uint32_t* lo = pre.lenoff + (uint64_t)tile * LZF_TILE;
                for (int r = tid; r < LZF_TILE; r += XPE_THREADS) sm.lz.lenoff[r] = lo[r];
                if (tid < LZF_TILE / 32) sm.lz.V[tid] = pre.visited[(uint64_t)tile * (LZF_TILE / 32) + tid];
                __syncthreads();
            } else if (EXACT) {
                const uint64_t ubase = (uint64_t)pre.seg_first[u] * XPE_SEG;
                const uint32_t* lo = pre.lenoff + ubase + t0;
                const uint32_t fe = sm.forced_end;
                for (int r = tid; r < LZF_TILE; r += XPE_THREADS) {
                    const uint32_t p = t0 + (uint32_t)r;
                    sm.lz.lenoff[r] = (p < n && p >= fe) ? lo[r] : 0u;       // (LZX_UNKNOWN where nobody has asked)
                }
                const uint64_t gt = (ubase + t0) / LZF_TILE;
                if (tid < LZF_TILE / 32) sm.lz.V[tid] = pre.visited[gt * (LZF_TILE / 32) + tid];
                if (tid == 0) sm.big_pos = LZF_TILE;
                __syncthreads();
                lzx_merge_parse(sm.lz, n, t0, pre.next_p[gt], fe, LzxUnit(src, n, pre.dist + ubase, 8192u, 0u), t0);
                // a visited match of LZX_CAP bytes or more leaves the tile, so there is at most one: measure it
                for (int r = tid; r < LZF_TILE; r += XPE_THREADS)
                    if (((sm.lz.V[r >> 5] >> (r & 31)) & 1) && (sm.lz.lenoff[r] >> 16) == LZX_CAP - 1u) sm.big_pos = (uint32_t)r;
                __syncthreads();
                const uint32_t bp = sm.big_pos;
                if (bp < (uint32_t)LZF_TILE) {
                    const uint32_t p = t0 + bp;
                    const uint32_t len = lzx_extend(src, p, sm.lz.lenoff[bp] & 0xFFFFu, n - p - 1u, &sm.red);
                    if (tid == 0) { sm.big_len = len; sm.lz.next_p = p + len; sm.forced_end = p + len + lzx_forced_literals(p, len, n); }
                    __syncthreads();
                }
            } else {
                lzf_tile(sm.lz, src, n, t0, 8192u, 8192u);
            }
            const uint32_t big_pos = EXACT ? sm.big_pos : (uint32_t)LZF_TILE, big_len = EXACT ? sm.big_len : 0u;
            auto len_of = [&](int r, uint32_t lo) -> uint32_t { return (EXACT && (uint32_t)r == big_pos) ? big_len : (lo >> 16) + 1u; };
            // ---- per-thread item statistics over 8 consecutive positions ----
            const uint32_t items0 = sm.items, bytes0 = sm.bytes, nibs0 = sm.nibs;
            uint32_t my_items = 0, my_bytes = 0, my_nibs = 0;
#pragma unroll
            for (int i = 0; i < XPE_PER_THREAD; ++i) {
                const int r = tid * XPE_PER_THREAD + i;
                if ((sm.lz.V[r >> 5] >> (r & 31)) & 1) {
                    const uint32_t lo = sm.lz.lenoff[r];
                    ++my_items;
                    if (lo && len_of(r, lo) - 3 >= 7) ++my_nibs;
                }
            }
            // The half-byte allocation depends on the GLOBAL parity of the nibble-user rank: scan the nibble users first,
            // then size the items with the right parity, then scan bytes and items.
            uint32_t nincl = (uint32_t)warp_incl_scan((int)my_nibs);
            if (lane == 31) sm.wsum[warp] = nincl;
            __syncthreads();
            uint32_t nbase = nibs0;
            for (int w2 = 0; w2 < warp; ++w2) nbase += (uint32_t)sm.wsum[w2];
            const uint32_t nib_excl = nbase + nincl - my_nibs;
            __syncthreads();
            {
                uint32_t q = nib_excl;
#pragma unroll
                for (int i = 0; i < XPE_PER_THREAD; ++i) {
                    const int r = tid * XPE_PER_THREAD + i;
                    if ((sm.lz.V[r >> 5] >> (r & 31)) & 1) {
                        const uint32_t lo = sm.lz.lenoff[r];
                        if (lo) {
                            const uint32_t lm3 = len_of(r, lo) - 3;
                            const bool nib = lm3 >= 7;
                            my_bytes += xpe_match_bytes(lm3, nib && ((q & 1) == 0));
                            if (nib) ++q;
                        } else my_bytes += 1;
                    }
                }
            }
            unsigned long long v = ((unsigned long long)my_bytes << 32) | my_items, incl = v;
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long x = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += x; }
            if (lane == 31) sm.wsum[warp] = incl;
            __syncthreads();
            unsigned long long base = 0, total = 0;
            for (int w2 = 0; w2 < XPE_THREADS / 32; ++w2) { if (w2 < warp) base += sm.wsum[w2]; total += sm.wsum[w2]; }
            const unsigned long long excl = base + incl - v;
            const uint32_t tile_items = (uint32_t)total, tile_bytes = (uint32_t)(total >> 32);
            // capacity: everything this tile can touch, including the flag word it may open at its end
            const uint32_t need = bytes0 + tile_bytes + 4 * ((items0 + tile_items) / 32 + 1);
            if (need > cap) { if (tid == 0) sm.status = ST_BUF; __syncthreads(); break; }

            // ---- flag words of this tile ----
            const uint32_t w_first = items0 >> 5;
            for (int i = tid; i < LZF_TILE / 32 + 2; i += XPE_THREADS) { sm.flagw[i] = (i == 0) ? sm.cur_flags : 0; sm.flagpos[i] = (i == 0) ? sm.cur_flagpos : 0; }
            __syncthreads();
            // ---- emit ----
            {
                uint32_t k = items0 + (uint32_t)excl, b = bytes0 + (uint32_t)(excl >> 32), q = nib_excl;
#pragma unroll
                for (int i = 0; i < XPE_PER_THREAD; ++i) {
                    const int r = tid * XPE_PER_THREAD + i;
                    if ((sm.lz.V[r >> 5] >> (r & 31)) & 1) {
                        const uint32_t lo = sm.lz.lenoff[r];
                        uint32_t pos = b + 4 * ((k >> 5) + 1);
                        if ((k & 31) == 0) sm.flagpos[(k >> 5) - w_first] = pos - 4;
                        if (lo) {
                            const uint32_t lm3 = len_of(r, lo) - 3, off = lo & 0xFFFFu;
                            const uint32_t sym = ((off - 1) << 3) | min(lm3, 7u);
                            atomicOr(&sm.flagw[(k >> 5) - w_first], 0x80000000u >> (k & 31));
                            o[pos] = (uint8_t)sym; o[pos + 1] = (uint8_t)(sym >> 8);
                            uint32_t e = pos + 2, sz = 2;
                            if (lm3 >= 7) {
                                const uint32_t nv = min(lm3 - 7, 15u);
                                if ((q & 1) == 0) { o[e] = (uint8_t)nv; sm.nibpos[q - nibs0] = e; ++e; ++sz; }
                                else sm.nibpos[q - nibs0] = 0x80000000u | nv;          // to be OR-ed into the previous user's byte
                                ++q;
                                if (lm3 - 7 >= 15) {
                                    o[e++] = (uint8_t)min(lm3 - 22, 255u); ++sz;
                                    if (lm3 - 22 >= 255) {
                                        if (lm3 <= 0xFFFFu) { o[e] = (uint8_t)lm3; o[e + 1] = (uint8_t)(lm3 >> 8); sz += 2; }
                                        else {                                   // src/xpress_compress.cpp:307-312
                                            o[e] = 0; o[e + 1] = 0; o[e + 2] = (uint8_t)lm3; o[e + 3] = (uint8_t)(lm3 >> 8);
                                            o[e + 4] = (uint8_t)(lm3 >> 16); o[e + 5] = (uint8_t)(lm3 >> 24); sz += 6;
                                        }
                                    }
                                }
                            }
                            b += sz;
                        } else { o[pos] = src[t0 + r]; b += 1; }
                        ++k;
                    }
                }
            }
            __syncthreads();
            // nibble users of the tile (= last thread's inclusive count); the odd-ranked ones complete the previous byte
            if (tid == XPE_THREADS - 1) sm.wsum[0] = (unsigned long long)(nib_excl + my_nibs - nibs0);
            __syncthreads();
            const uint32_t tile_nibs = (uint32_t)sm.wsum[0];
            for (uint32_t j = tid; j < tile_nibs; j += XPE_THREADS) {
                const uint32_t e = sm.nibpos[j];
                if (e & 0x80000000u) {                                     // odd-ranked user
                    const uint32_t where = (j == 0) ? sm.pend_half : sm.nibpos[j - 1];
                    o[where] = (uint8_t)(o[where] | ((e & 15u) << 4));
                }
            }
            // complete flag words go out; the last (possibly partial) one is carried
            const uint32_t w_last = (items0 + tile_items) >> 5;            // word that receives the next item
            for (uint32_t w = w_first + tid; w < w_last; w += XPE_THREADS) {
                const uint32_t fv = sm.flagw[w - w_first], fp = sm.flagpos[w - w_first];
                o[fp] = (uint8_t)fv; o[fp + 1] = (uint8_t)(fv >> 8); o[fp + 2] = (uint8_t)(fv >> 16); o[fp + 3] = (uint8_t)(fv >> 24);
            }
            __syncthreads();
            if (tid == 0) {
                if (tile_nibs) {
                    const uint32_t lastq = tile_nibs - 1;
                    if (!(sm.nibpos[lastq] & 0x80000000u)) sm.pend_half = sm.nibpos[lastq];
                }
                sm.items = items0 + tile_items; sm.bytes = bytes0 + tile_bytes; sm.nibs = nibs0 + tile_nibs;
                if (((items0 + tile_items) & 31) == 0) {
                    // the word for the next 32 items is reserved right after the last item (:317-324)
                    sm.cur_flags = 0; sm.cur_flagpos = bytes0 + tile_bytes + 4 * (w_last);
                } else { sm.cur_flags = sm.flagw[w_last - w_first]; sm.cur_flagpos = sm.flagpos[w_last - w_first]; }
            }
            __syncthreads();
        }
        if (tid == 0) {
            uint32_t total = 0;
            if (sm.status == ST_OK) {
                const uint32_t cnt = sm.items & 31;
                uint32_t fv, fp;
                if (n == 0) { fv = 0xFFFFFFFFu; fp = 0; total = 4; }
                else {
                    fp = sm.cur_flagpos;
                    fv = cnt ? (sm.cur_flags | (0xFFFFFFFFu >> cnt)) : 0xFFFFFFFFu;     // :343
                    total = sm.bytes + 4 * ((sm.items >> 5) + 1);
                }
                if (total > cap) sm.status = ST_BUF;
                else { o[fp] = (uint8_t)fv; o[fp + 1] = (uint8_t)(fv >> 8); o[fp + 2] = (uint8_t)(fv >> 16); o[fp + 3] = (uint8_t)(fv >> 24); }
            }
            out_len[u] = sm.status == ST_OK ? total : 0;
            status[u] = sm.status;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(XPE_THREADS)
xpress_encode_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                     const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                     const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                     uint32_t* __restrict__ out_len, int32_t* __restrict__ status,
                     uint32_t seg_tiles, XpressSeg* __restrict__ index, uint32_t index_stride)     // optional restart index
{
    xpress_encode_body<0>(in, out, count, in_off, in_len, out_off, out_cap, out_len, status, seg_tiles, index, index_stride, XpPre{});
}

// Pass 2 of the two-pass variant: emission from the per-tile results of xpress_match_kernel.
__global__ void __launch_bounds__(XPE_THREADS)
xpress_emit_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                   const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                   const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                   uint32_t* __restrict__ out_len, int32_t* __restrict__ status,
                   uint32_t seg_tiles, XpressSeg* __restrict__ index, uint32_t index_stride, XpPre pre)
{
    xpress_encode_body<1>(in, out, count, in_off, in_len, out_off, out_cap, out_len, status, seg_tiles, index, index_stride, pre);
}

// Emission of the reference's parse: pre.lenoff = the answers of lz_find_kernel (indexed seg_first[u] * XPE_SEG + position),
// pre.visited / pre.next_p = its per-tile parse from the tile's first position (visited0 / exit0).
__global__ void __launch_bounds__(XPE_THREADS)
xpress_emit_exact_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                         const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                         const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                         uint32_t* __restrict__ out_len, int32_t* __restrict__ status,
                         uint32_t seg_tiles, XpressSeg* __restrict__ index, uint32_t index_stride, XpPre pre)
{
    xpress_encode_body<2>(in, out, count, in_off, in_len, out_off, out_cap, out_len, status, seg_tiles, index, index_stride, pre);
}

// seg_first[u] = segments of the buffers before u, seg_first[count] = all of them; totals[0] = that, totals[1] = input bytes,
// totals[2] = the longest buffer.
// Single block; also clears the four work counters of the kernels that follow.
__global__ void __launch_bounds__(1024)
xpress_segments_kernel(const uint32_t* __restrict__ in_len, uint32_t count, uint32_t* __restrict__ seg_first, unsigned long long* __restrict__ totals,
                       uint32_t* __restrict__ work_counter)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    __shared__ unsigned long long bytes;
    __shared__ uint32_t longest;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; bytes = 0; longest = 0; }
    if (tid < 4) work_counter[tid] = 0;                                  // one counter per persistent kernel of the pipeline
    __syncthreads();
    unsigned long long my_bytes = 0;
    uint32_t my_longest = 0;
    for (uint32_t u0 = 0; u0 < count; u0 += 1024) {
        const uint32_t u = u0 + (uint32_t)tid;
        const uint32_t n = u < count ? in_len[u] : 0u;
        const uint32_t v = u < count ? (n + XPE_SEG - 1) / XPE_SEG : 0u;      // (an empty buffer has no segment: the emitter handles it alone)
        my_bytes += n;
        my_longest = max(my_longest, n);
        const uint32_t incl = (uint32_t)warp_incl_scan((int)v);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t woff = carry;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        if (u < count) seg_first[u] = woff + incl - v;
        __syncthreads();
        if (tid == 1023) carry = woff + incl;
        __syncthreads();
    }
    atomicAdd(&bytes, my_bytes);
    atomicMax(&longest, my_longest);
    __syncthreads();
    if (tid == 0) { seg_first[count] = carry; totals[0] = carry; totals[1] = bytes; totals[2] = longest; }
}

// Pass 1 of the two-pass variant: persistent blocks pull (buffer, segment) pairs from a work counter and run the LZ77 front end
// over the segment as if it were a buffer of its own.
__global__ void __launch_bounds__(XPE_THREADS)
xpress_match_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                    const uint32_t* __restrict__ seg_first, uint32_t* __restrict__ work_counter,
                    uint32_t* __restrict__ lenoff, uint32_t* __restrict__ visited, uint32_t* __restrict__ next_p)
{
    MSC_DYN_SMEM(smem_raw);
    LzFrontSmem& lz = *reinterpret_cast<LzFrontSmem*>(smem_raw);
    __shared__ uint32_t s_work;
    const int tid = (int)threadIdx.x;
    const uint32_t total = seg_first[count];
    for (;;) {
        if (tid == 0) s_work = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t w = s_work;
        __syncthreads();
        if (w >= total) break;
        uint32_t lo = 0, hi = count;                                  // the buffer u with seg_first[u] <= w < seg_first[u + 1]
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seg_first[mid] <= w) lo = mid; else hi = mid; }
        const uint32_t u = lo, sg = w - seg_first[u];
        const uint8_t* src = in + in_off[u];                          // positions are relative to the buffer: matches may reach back
        const uint32_t seg_start = sg * XPE_SEG;                      // into the previous segment (that is earlier data of the same stream)
        const uint32_t n = min(seg_start + XPE_SEG, in_len[u]);       // ... but neither a match nor the parse crosses the segment's end
        for (int i = tid; i < (1 << LZF_HBITS); i += XPE_THREADS) lz.H[i] = 0;
        __syncthreads();
        if (sg) lzf_warm(lz, src, n, seg_start - 8192u, seg_start);   // the window's worth of history before the segment
        if (tid == 0) { lz.next_p = seg_start; lz.ntok = 0; }
        __syncthreads();
        for (uint32_t t0 = seg_start; t0 < n; t0 += LZF_TILE) {
            const uint64_t tile = (uint64_t)w * XPE_SEG_TILES + (t0 - seg_start) / LZF_TILE;
            if (tid == 0) next_p[tile] = lz.next_p - seg_start;
            lzf_tile(lz, src, n, t0, 8192u, 8192u);
            for (int r = tid; r < LZF_TILE; r += XPE_THREADS) lenoff[tile * LZF_TILE + r] = lz.lenoff[r];
            if (tid < LZF_TILE / 32) visited[tile * (LZF_TILE / 32) + tid] = lz.V[tid];
            __syncthreads();
        }
    }
}

} // namespace msc
