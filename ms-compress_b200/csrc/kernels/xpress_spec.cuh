// xpress_spec.cuh -- ONE plain-Xpress stream decoded on the whole GPU: the flag groups are found by speculation.
//
// The reference's xpress_decompress (src/xpress_decompress.cpp:405-462) is one loop over the stream: a 32-bit flag word, then 32
// items (a literal byte, or a 16-bit match symbol with optional length bytes and a length nibble that two matches share).  The
// stream carries no marker to start a second reader from, so the one-shot call on a single stream used to be one
// warp.  But the reader's state at a flag word is tiny:
//
//     (p, c)    p = offset of the flag word,  c = 0: no half byte pending, 1: a half byte pending whose high nibble is < 15,
//                                                 2: pending and the nibble is 15 (then length bytes follow in the stream)
//
// -- the position of the pending half byte only matters for the VALUE of a length, never for where the next item starts.  So:
//   groups  : for EVERY offset p and every c the group that would start there is walked once (xps_groups_kernel): where the next
//             flag word lies, the next c, the bytes the group produces, the half byte it leaves -- one 64-bit record per state;
//   blocks  : XPS_BLK offsets form a block; pointer doubling in shared memory composes the records of a block into "entered in
//             state e, the reader leaves the block in state x with half byte L" for the few states a block can be entered in
//             (a group is at most 324 bytes long, so an entry lies in the first 324 offsets) (xps_blocks_kernel);
//   entries : one thread hops from block to block, one table look-up each, from state (0, 0) (xps_entries_kernel);
//   walk    : with its true entry state a block walks its own true groups -- count and bytes first, then, after a scan over the
//             blocks, a restart entry (XpressSeg: in_pos, out_pos, half_pos) at every XPS_EVERY-th group (xps_walk_kernel);
//   after that the stream has a restart index like the ones this library's encoder writes, and the indexed parser
//   (xpress_parse_indexed_kernel, one lane per segment) and the order-free token execution (lz_jump.cuh) do the rest.
// Nothing the speculation says is trusted: the indexed parser re-reads every item with the reference's checks, every segment has to
// end exactly where the next begins (xps_chain_kernel), and anything else -- a damaged stream, too little room -- is left to the
// serial decoder, which owns the status codes.
#pragma once
#include "common.cuh"
#include "xpress_encode.cuh"      // XpressSeg

namespace msc {

constexpr uint32_t XPS_BLK = 8192;                    // offsets per block
constexpr uint32_t XPS_NST = XPS_BLK * 3;             // states of a block
constexpr uint32_t XPS_MAXGROUP = 4 + 32 * 10;        // longest group: flag word + 32 x (symbol 2, nibble byte 1, length 1 + 2 + 4)
constexpr uint32_t XPS_NENT = (XPS_MAXGROUP + 1) * 3; // states a block can be entered in / left to (offsets 0..324 past the block's end)
constexpr uint32_t XPS_TERM = 0xFFFEu, XPS_BAD = 0xFFFFu, XPS_SAME = 0xFFFFu;
constexpr int XPS_THREADS = 1024;
constexpr int XPS_PER_THREAD = (int)(XPS_NST / XPS_THREADS);
constexpr int XPS_ROUNDS = 8;                         // a full group is at least 36 bytes: at most 228 groups per block < 2^8
constexpr uint32_t XPS_EVERY = 16;                    // one restart entry per this many groups
constexpr uint32_t XPS_NONE = 0xFFFFFFFFu;
static_assert(XPS_NST % XPS_THREADS == 0, "states per thread");
static_assert(XPS_NST + XPS_NENT < XPS_TERM, "local state numbers fit 16 bits");

// record of one state: dp | c' << 9 | kind << 11 | used << 13 | hl << 14 | bytes << 32
//   dp   next flag word = p + dp;  c' state there;  kind 0: go on, 1: the stream ends inside this group, 2: no valid group here
//   used the group took the pending nibble of a c == 1 state: its value (0..14) is NOT in `bytes`
//   hl   != 0: the group leaves a new half byte pending at p + hl;  0 with c' != 0: the one it came with is still pending
__device__ __forceinline__ uint32_t xps_dp(unsigned long long r) { return (uint32_t)r & 0x1FFu; }
__device__ __forceinline__ uint32_t xps_c(unsigned long long r) { return ((uint32_t)r >> 9) & 3u; }
__device__ __forceinline__ uint32_t xps_kind(unsigned long long r) { return ((uint32_t)r >> 11) & 3u; }
__device__ __forceinline__ uint32_t xps_used(unsigned long long r) { return ((uint32_t)r >> 13) & 1u; }
__device__ __forceinline__ uint32_t xps_hl(unsigned long long r) { return ((uint32_t)r >> 14) & 0x1FFu; }
__device__ __forceinline__ uint32_t xps_bytes(unsigned long long r) { return (uint32_t)(r >> 32); }

// The group at offset p (p + 4 <= n) read in state c: src/xpress_decompress.cpp:425-459 without the output.
__device__ __forceinline__ unsigned long long xps_group(const uint8_t* __restrict__ in, uint32_t n, uint32_t p, uint32_t c)
{
    const uint32_t flags = (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
    uint32_t ip = p + 4;
    bool has_half = c != 0, incoming = c != 0;
    uint32_t half_loc = 0, used = 0, kind = 0;
    unsigned long long bytes = 0;
    for (int k = 0; k < 32; ++k) {
        if (ip == n) { kind = 1; break; }                                         // :433-438 (the indexed parser checks the flags)
        if ((flags >> (31 - k)) & 1u) {
            if (ip + 2 > n) { kind = 2; break; }
            uint32_t len = in[ip] & 7u;
            ip += 2;
            if (len == 7) {
                if (has_half) {
                    if (incoming) { len = c == 2 ? 15u : 0u; used = c == 1 ? 1u : 0u; }
                    else len = in[half_loc] >> 4;
                    has_half = false;
                } else if (ip == n) { kind = 2; break; }
                else { half_loc = ip; has_half = true; incoming = false; len = in[ip++] & 0xFu; }
                if (len == 15) {
                    if (ip == n) { kind = 2; break; }
                    len = in[ip++];
                    if (len == 0xFF) {
                        if (ip + 2 > n) { kind = 2; break; }
                        len = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8); ip += 2;
                        if (len == 0) {
                            if (ip + 4 > n) { kind = 2; break; }
                            len = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8) | ((uint32_t)in[ip + 2] << 16) | ((uint32_t)in[ip + 3] << 24); ip += 4;
                        }
                        if (len < 0xF + 0x7) { kind = 2; break; }
                        len -= 0xF + 0x7;
                    }
                    bytes += 0xF;
                }
                bytes += 0x7;
            }
            bytes += (unsigned long long)len + 3;
        } else { bytes += 1; ip += 1; }
    }
    if (kind == 0 && ip + 4 > n) kind = 2;                                        // :461 the stream may only end inside a group
    if (bytes > 0x7FFFFFFFull) kind = 2;                                          // (not an error, but nothing this path handles)
    uint32_t c2 = 0, hl = 0;
    if (has_half) {
        if (incoming) c2 = c;
        else { c2 = (in[half_loc] >> 4) == 15 ? 2u : 1u; hl = half_loc - p; }
    }
    return (unsigned long long)(ip - p) | (c2 << 9) | (kind << 11) | (used << 13) | (hl << 14) | (bytes << 32);
}

// rec[p * 3 + c] for every offset p with p + 4 <= n.
__global__ void __launch_bounds__(256)
xps_groups_kernel(const uint8_t* __restrict__ in, uint32_t n, unsigned long long* __restrict__ rec)
{
    const uint32_t npos = n - 3;
    const uint64_t nstate = (uint64_t)npos * 3;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nstate; i += (uint64_t)gridDim.x * blockDim.x) {
        // consecutive threads take consecutive offsets (their reads overlap); the three states of an offset are a third of the grid apart
        const uint32_t c = (uint32_t)(i / npos), p = (uint32_t)(i % npos);
        rec[(uint64_t)p * 3 + c] = xps_group(in, n, p, c);
    }
}

// One block of XPS_BLK offsets: table[b * XPS_NENT + e] = (state the reader leaves the block in) | (half byte it leaves pending) << 16
// for the entry state e = (offset - block start) * 3 + c.  Leaving states are numbered from the next block's start, so they ARE the
// next block's entry numbers; XPS_TERM: the stream ends inside the block; XPS_BAD: no valid parse.  Half byte: offset - block start,
// XPS_SAME = the one the reader came with (or none).
__global__ void __launch_bounds__(XPS_THREADS)
xps_blocks_kernel(const unsigned long long* __restrict__ rec, uint32_t n, uint32_t nblk, uint32_t* __restrict__ table)
{
    MSC_DYN_SMEM(smem_raw);
    uint32_t* T = reinterpret_cast<uint32_t*>(smem_raw);                          // [XPS_NST]
    const uint32_t npos = n - 3;
    const int tid = (int)threadIdx.x;
    for (uint32_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const uint32_t p0 = b * XPS_BLK;
        for (uint32_t s = (uint32_t)tid; s < XPS_NST; s += XPS_THREADS) {
            const uint32_t p = p0 + s / 3;
            uint32_t e = XPS_BAD | (XPS_SAME << 16);
            if (p < npos) {
                const unsigned long long r = rec[(uint64_t)p0 * 3 + s];
                const uint32_t kind = xps_kind(r), c2 = xps_c(r), hl = xps_hl(r);
                const uint32_t to = p + xps_dp(r) - p0;                           // next flag word, relative to the block
                const uint32_t id = kind == 2 ? XPS_BAD : kind == 1 ? XPS_TERM : to * 3 + c2;   // (to >= XPS_BLK: a leaving state)
                const uint32_t L = (c2 != 0 && hl != 0) ? (p + hl - p0) : XPS_SAME;
                e = id | (L << 16);
            }
            T[s] = e;
        }
        __syncthreads();
        for (int r = 0; r < XPS_ROUNDS; ++r) {
            uint32_t v[XPS_PER_THREAD];
#pragma unroll
            for (int k = 0; k < XPS_PER_THREAD; ++k) {
                const uint32_t e = T[tid + k * XPS_THREADS];
                const uint32_t id = e & 0xFFFFu;
                uint32_t o = e;
                if (id < XPS_NST) {
                    const uint32_t e2 = T[id];
                    o = (e2 & 0xFFFFu) | (((e2 >> 16) != XPS_SAME) ? (e2 & 0xFFFF0000u) : (e & 0xFFFF0000u));
                }
                v[k] = o;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < XPS_PER_THREAD; ++k) T[tid + k * XPS_THREADS] = v[k];
            __syncthreads();
        }
        for (uint32_t e = (uint32_t)tid; e < XPS_NENT; e += XPS_THREADS) {
            uint32_t v = T[e];
            const uint32_t id = v & 0xFFFFu;
            if (id < XPS_NST) v = XPS_BAD | (XPS_SAME << 16);                      // (cannot happen: XPS_ROUNDS doublings cover a block)
            else if (id < XPS_TERM) v = (id - XPS_NST) | (v & 0xFFFF0000u);
            table[(uint64_t)b * XPS_NENT + e] = v;
        }
        __syncthreads();
    }
}

struct XpsSummary {
    uint32_t ok;              // entries: the chain of blocks reaches the end of the stream; scan: ... and the sizes fit
    uint32_t last_blk;        // block the stream ends in
    uint32_t nseg;            // restart entries
    uint32_t groups;          // groups of the stream
    unsigned long long bytes; // bytes the stream produces
};

// From state (0, 0) block by block: entry[b] = entry state number of block b, entry_half[b] = offset of the half byte pending there.
__global__ void __launch_bounds__(32)
xps_entries_kernel(const uint32_t* __restrict__ table, uint32_t nblk, uint32_t* __restrict__ entry, uint32_t* __restrict__ entry_half,
                   XpsSummary* __restrict__ sum)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t e = 0, L = XPS_NONE, ok = 0, last = 0;
    for (uint32_t b = 0; b < nblk; ++b) {
        entry[b] = e; entry_half[b] = L;
        const uint32_t v = table[(uint64_t)b * XPS_NENT + e];
        const uint32_t id = v & 0xFFFFu, l = v >> 16;
        if (l != XPS_SAME) L = b * XPS_BLK + l;
        if (id == XPS_TERM) { ok = 1; last = b; break; }
        if (id >= XPS_NENT) break;                                                // XPS_BAD
        e = id;
    }
    sum->ok = ok; sum->last_blk = last; sum->nseg = 0; sum->groups = 0; sum->bytes = 0;
}

// The true groups of block b, from its true entry state.  EMIT = false: count[b] = groups << 40 | bytes.  EMIT = true: base[b] holds
// the same two numbers summed over the blocks before b; every XPS_EVERY-th group of the stream becomes a restart entry.
template <bool EMIT>
__global__ void __launch_bounds__(32)
xps_walk_kernel(const uint8_t* __restrict__ in, uint32_t n, const unsigned long long* __restrict__ rec, const uint32_t* __restrict__ entry,
                const uint32_t* __restrict__ entry_half, const XpsSummary* __restrict__ sum, unsigned long long* __restrict__ count,
                const unsigned long long* __restrict__ base, XpressSeg* __restrict__ index)
{
    if (threadIdx.x != 0 || !sum->ok) return;
    const uint32_t nb = sum->last_blk + 1;
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const uint32_t e = entry[b], end = (b + 1) * XPS_BLK;
        uint32_t p = b * XPS_BLK + e / 3, c = e % 3, L = entry_half[b];
        unsigned long long g = 0, o = 0;
        if (EMIT) { g = base[b] >> 40; o = base[b] & 0xFFFFFFFFFFull; }
        const unsigned long long g0 = g, o0 = o;
        for (;;) {
            const unsigned long long r = rec[(uint64_t)p * 3 + c];
            if (EMIT && g % XPS_EVERY == 0) {
                XpressSeg s;
                s.in_pos = p; s.out_pos = (uint32_t)o; s.flag_pos = 0; s.flag_used = 0; s.half_pos = c ? L : XPS_NONE;
                s.nseg = g == 0 ? sum->nseg : 0;
                index[g / XPS_EVERY] = s;
            }
            ++g;
            o += xps_bytes(r) + ((xps_used(r) && L < n) ? (uint32_t)(in[L] >> 4) : 0u);
            const uint32_t c2 = xps_c(r), hl = xps_hl(r);
            if (c2 == 0) L = XPS_NONE; else if (hl) L = p + hl;
            if (xps_kind(r) != 0) break;                                          // the last group of the stream (kind 2 is not on a true path)
            p += xps_dp(r); c = c2;
            if (p >= end) break;
        }
        if (!EMIT) count[b] = ((g - g0) << 40) | (o - o0);
    }
}

// Exclusive sums of count[] over the blocks of the stream, the totals and the number of restart entries.
__global__ void __launch_bounds__(32)
xps_scan_kernel(const unsigned long long* __restrict__ count, unsigned long long* __restrict__ base, XpsSummary* __restrict__ sum,
                unsigned long long cap)
{
    if (blockIdx.x != 0 || threadIdx.x != 0 || !sum->ok) return;
    const uint32_t nb = sum->last_blk + 1;
    unsigned long long g = 0, o = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        base[b] = (g << 40) | o;
        const unsigned long long c = count[b];
        g += c >> 40; o += c & 0xFFFFFFFFFFull;
        if (o > 0x7FFFFFFFull || g >= (1ull << 24)) { sum->ok = 0; return; }
    }
    if (o > cap) { sum->ok = 0; return; }                                        // MSCOMP_BUF_ERROR is the serial decoder's to report
    sum->groups = (uint32_t)g; sum->bytes = o; sum->nseg = (uint32_t)((g + XPS_EVERY - 1) / XPS_EVERY);
}

// After the indexed parse: every segment parsed cleanly and ends where the next one starts?  Then chain[] = slot[nseg] (u64) |
// ntok[nseg] | pos[nseg] as the token executors take them (lz_jump.cuh), and result = {ok, bytes}.
__global__ void __launch_bounds__(256)
xps_chain_kernel(const XpressSeg* __restrict__ index, uint32_t nseg, const uint32_t* __restrict__ seg_ntok, const int32_t* __restrict__ seg_status,
                 const uint32_t* __restrict__ seg_end, unsigned long long* __restrict__ slot, uint32_t* __restrict__ ntok, uint32_t* __restrict__ pos,
                 uint32_t* __restrict__ result)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nseg; j += gridDim.x * blockDim.x) {
        const uint32_t at = index[j].out_pos;
        slot[j] = at; ntok[j] = seg_ntok[j]; pos[j] = at;
        const bool good = seg_status[j] == ST_OK && (j + 1 < nseg ? seg_end[j] == index[j + 1].out_pos : true);
        if (!good) atomicAdd(&result[0], 1u);
        if (j + 1 == nseg) result[1] = seg_end[j];
    }
}

} // namespace msc
