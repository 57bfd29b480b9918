// lznt1_decode.cuh -- LZNT1 chunk decoder, one warp per 4 KiB chunk.
//
// Replaces the reference's lznt1_decompress_chunk (src/lznt1_decompress.cpp:37-121) and the per-chunk part of
// lznt1_decompress_chunk_read (:122-209: header parse, signature check, stored-chunk copy).
//
// Wire format: u16 header = (payload bytes - 1) | 0x3000 | 0x8000 if compressed; compressed
// payload = { flag byte (LSB first, 1 = token); up to 8 items: literal byte or u16 token }.  A token at chunk
// output position p splits as off = (tok >> shift) + 1, len = (tok & ((1 << shift) - 1)) + 3 with
// shift = 12 while p <= 16 and one less each time p passes a power of two (reference :59-63).
//
// Parallel scheme (per warp, chunk staged in shared memory):
//   1. "chase": the start of flag group g+1 is start(g) + 9 + popc(flags(g)) -- a serial chain that does not
//      depend on token contents; all lanes walk it redundantly, lane i keeps group i (<= 32 groups per round).
//   2. every lane sums the output length of its own group assuming the round's shift is the one in force at
//      the round's first byte; a warp scan turns that into output positions; groups up to the first one that
//      has a token starting beyond the current power-of-two limit are exact ("valid").
//   3. valid lanes emit: literals as they walk their items; matches under a watermark rule -- a pending match
//      may run when every byte it reads lies below the lowest pending match destination of the whole warp
//      (everything below that is final), or when it is that lowest one itself.  Long matches are copied by
//      the whole warp.
//   4. the first invalid group (at most one per power of two, <= 9 per chunk) is decoded item by item with the
//      exact shift ("careful" path), warp-uniform.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int LZNT1_CHUNK = 4096;
constexpr int LZNT1_IN_SMEM = 4096 + 32;          // payload (<= 4096) + alignment skew (<= 15) rounded up to 16
constexpr int LZNT1_WARPS_PER_BLOCK = 4;
constexpr int LZNT1_COOP_LEN = 24;                // matches at least this long are copied by the whole warp

struct Lznt1WarpSmem {
    __align__(16) uint8_t in[LZNT1_IN_SMEM];
    __align__(16) uint8_t out[LZNT1_CHUNK];
};

// shift / limit in force for a token that starts at chunk output position p
__device__ __forceinline__ void lznt1_shift_limit(int p, int& shift, int& limit)
{
    const int m = (p <= 16) ? 4 : (32 - __clz(p - 1));
    shift = 16 - m;
    limit = 1 << m;
}

// whole-warp copy of a match inside the shared output buffer (handles overlap like a byte-serial copy)
__device__ __forceinline__ void lznt1_warp_match_copy(uint8_t* out, int dst, int off, int len)
{
    const int lane = lane_id();
    if (off >= 32) {
        for (int i = 0; i < len; i += 32) {
            if (i + lane < len) out[dst + i + lane] = out[dst - off + i + lane];
            __syncwarp();
        }
    } else {
        // period `off` pattern: every byte equals one of the `off` bytes just before dst (all final)
        for (int i = lane; i < len; i += 32) out[dst + i] = out[dst - off + (i % off)];
        __syncwarp();
    }
}

// Decode one compressed payload s_in[0..n) into s_out; returns bytes produced or a negative status.
__device__ __forceinline__ int lznt1_decode_payload(const uint8_t* __restrict__ s_in, int n, uint8_t* __restrict__ s_out)
{
    const int lane = lane_id();
    int q = 0;          // payload offset of the next flag byte (uniform)
    int P = 0;          // output bytes produced so far (uniform)
    int err = 0;        // uniform

    while (q < n && !err) {
        int S, lim;
        lznt1_shift_limit(P, S, lim);
        const int mask = (1 << S) - 1;
        int R = ((lim - P) >> 3) + 2;              // no more groups than this can be valid in this round
        if (R > 32) R = 32;

        // ---- 1. chase the group starts -------------------------------------------------------------
        int s = q, my = -1;
        unsigned myf = 0;
        for (int i = 0; i < R; ++i) {
            if (s < n) {
                const unsigned f = s_in[s];
                if (lane == i) { my = s; myf = f; }
                s += 9 + __popc(f);
            }
        }

        // ---- 2. group lengths under the round's shift ----------------------------------------------
        int glen = 0, last_tok = -1;
        bool trunc = false;                         // token cut by the end of the payload
        if (my >= 0) {
            int o = my + 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (o < n) {
                    if ((myf >> j) & 1) {
                        if (o + 2 > n) { trunc = true; o = n; }
                        else {
                            const int t = s_in[o] | (s_in[o + 1] << 8);
                            last_tok = glen;
                            glen += (t & mask) + 3;
                            o += 2;
                        }
                    } else { glen += 1; o += 1; }
                }
            }
        }
        const int incl = warp_incl_scan(glen);
        const int gs = P + incl - glen;             // my group's first output position (exact while all earlier lanes valid)
        const bool valid = (my >= 0) && (last_tok < 0 || gs + last_tok <= lim);
        const unsigned act = __ballot_sync(FULL, my >= 0);
        const unsigned inv = __ballot_sync(FULL, my >= 0 && !valid);
        const int nact = __popc(act);
        const int G = inv ? (__ffs((int)inv) - 1) : nact;      // number of leading valid groups
        // position after the valid groups
        const int Pend = (G > 0) ? (P + __shfl_sync(FULL, incl, G - 1)) : P;
        if (Pend > LZNT1_CHUNK) { err = ST_DATA; break; }
        if (__ballot_sync(FULL, trunc && lane < G)) { err = ST_DATA; break; }

        // ---- 3. emit the valid groups -----------------------------------------------------------------
        {
            const bool mine = lane < G;
            // 3a. literals: every lane walks its (at most 8) items once; positions follow from the group start
            unsigned mb = 0;                                   // my match items still to execute
            if (mine) {
                int o = my + 1, run = gs;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (o < n) {
                        if ((myf >> j) & 1) {
                            const int t = s_in[o] | (s_in[o + 1] << 8);
                            run += (t & mask) + 3;
                            o += 2;
                            mb |= 1u << j;
                        } else { s_out[run++] = s_in[o++]; }
                    }
                }
            }
            // 3b. matches, in position order per lane, under the watermark rule.  The next match of a lane is found
            //     with find-first-set; its input / output offsets follow from popcounts (tokens are 2 bytes, literals 1).
            int acc = 0;                                       // output bytes of my matches executed so far
            bool bad = false;
            for (;;) {
                int pos = 0, m_len = 0, m_off = 0;
                const bool pending = mb != 0;
                if (pending) {
                    const int j = __ffs((int)mb) - 1;
                    const int tb = __popc(myf & ((1u << j) - 1));          // tokens before item j
                    const int o = my + 1 + j + tb;
                    pos = gs + (j - tb) + acc;
                    const int t = s_in[o] | (s_in[o + 1] << 8);
                    m_len = (t & mask) + 3; m_off = (t >> S) + 1;
                    if (m_off > pos) { bad = true; mb = 0; }
                }
                __syncwarp();
                const bool live = pending && !bad;
                const unsigned W = __reduce_min_sync(FULL, live ? (unsigned)pos : 0xffffffffu);
                if (W == 0xffffffffu) break;
                const bool can = live && ((unsigned)pos == W || (unsigned)(pos - m_off + m_len) <= W);
                unsigned coop = __ballot_sync(FULL, can && m_len >= LZNT1_COOP_LEN);
                if (can && m_len < LZNT1_COOP_LEN) {
                    for (int i = 0; i < m_len; ++i) s_out[pos + i] = s_out[pos - m_off + i];
                }
                while (coop) {
                    const int src = __ffs((int)coop) - 1;
                    coop &= coop - 1;
                    const int d = __shfl_sync(FULL, pos, src), l = __shfl_sync(FULL, m_len, src), f = __shfl_sync(FULL, m_off, src);
                    lznt1_warp_match_copy(s_out, d, f, l);
                }
                if (can) { acc += m_len; mb &= mb - 1; }
            }
            if (__ballot_sync(FULL, bad)) { err = ST_DATA; break; }
        }
        P = Pend;

        // ---- 4. careful path for the first invalid group ------------------------------------------------
        if (G < nact) {
            const int cs = __shfl_sync(FULL, my, G);
            const unsigned cf = __shfl_sync(FULL, myf, G);
            int o = cs + 1;
            for (int j = 0; j < 8 && o < n; ++j) {
                if ((cf >> j) & 1) {
                    if (o + 2 > n) { err = ST_DATA; break; }
                    int S2, lim2;
                    lznt1_shift_limit(P, S2, lim2);
                    const int t = s_in[o] | (s_in[o + 1] << 8);
                    const int len = (t & ((1 << S2) - 1)) + 3, off = (t >> S2) + 1;
                    o += 2;
                    if (off > P || P + len > LZNT1_CHUNK) { err = ST_DATA; break; }
                    __syncwarp();
                    lznt1_warp_match_copy(s_out, P, off, len);
                    P += len;
                } else {
                    if (P >= LZNT1_CHUNK) { err = ST_DATA; break; }
                    if (lane == 0) s_out[P] = s_in[o];
                    ++P; ++o;
                }
            }
            __syncwarp();
            // next round starts after the careful group
            q = (G + 1 < nact) ? __shfl_sync(FULL, my, G + 1) : s;
        } else {
            q = s;
        }
    }
    __syncwarp();
    return err ? err : P;
}

// One warp per chunk.  chunk_off[c] = byte offset of chunk c's header inside `in`.
// Chunk c is written to out + c*4096 (slot layout); out_size[c] = bytes produced, status[c] = 0 or < 0.
__global__ void __launch_bounds__(LZNT1_WARPS_PER_BLOCK * 32)
lznt1_decode_kernel(const uint8_t* __restrict__ in, uint64_t in_len, const uint64_t* __restrict__ chunk_off, uint32_t nchunks,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ out_size, int32_t* __restrict__ status)
{
    __shared__ Lznt1WarpSmem smem[LZNT1_WARPS_PER_BLOCK];
    const int lane = lane_id();
    const int wib = (int)(threadIdx.x >> 5);
    Lznt1WarpSmem& sm = smem[wib];
    const uint32_t warps_total = gridDim.x * LZNT1_WARPS_PER_BLOCK;

    for (uint32_t c = blockIdx.x * LZNT1_WARPS_PER_BLOCK + wib; c < nchunks; c += warps_total) {
        const uint64_t start = chunk_off[c];
        uint8_t* dst = out + (uint64_t)c * LZNT1_CHUNK;
        int st = ST_OK;
        uint32_t produced = 0;
        if (start + 2 > in_len) { st = ST_BUF; }
        else {
            const uint32_t hdr = (uint32_t)in[start] | ((uint32_t)in[start + 1] << 8);
            const uint32_t n = (hdr & 0x0FFF) + 1;                       // payload bytes
            if (hdr == 0) { st = ST_DATA; }                              // terminator is handled by the caller's index walk
            else if (start + 2 + n > in_len) { st = ST_BUF; }            // truncated chunk (reference: internal.h:627)
            else if ((hdr & 0x7000) != 0x3000) { st = ST_DATA; }         // src/lznt1_decompress.cpp:152
            else if (!(hdr & 0x8000)) {                                  // stored chunk :188-206
                warp_copy_bytes(dst, in + start + 2, n);
                produced = n;
            } else {
                // stage the payload: 16-byte granules covering [src, src+n)
                const uint8_t* src = in + start + 2;
                const uint32_t skew = (uint32_t)((uintptr_t)src & 15u);
                const uint8_t* a0 = src - skew;
                const uint32_t wbytes = (skew + n + 15u) & ~15u;
                for (uint32_t i = lane * 16; i < wbytes; i += 512)
                    *reinterpret_cast<uint4*>(sm.in + i) = ld_stream16(a0 + i);
                __syncwarp();
                const int r = lznt1_decode_payload(sm.in + skew, (int)n, sm.out);
                if (r < 0) st = r;
                else {
                    produced = (uint32_t)r;
                    // shared -> global, 16-byte vectors (dst is 16-byte aligned when `out` is)
                    const uint32_t nvec = produced >> 4;
                    if (((uintptr_t)dst & 15u) == 0) {
                        for (uint32_t v = lane; v < nvec; v += 32)
                            st_stream16(dst + 16 * v, *reinterpret_cast<const uint4*>(sm.out + 16 * v));
                        for (uint32_t b = (nvec << 4) + lane; b < produced; b += 32) dst[b] = sm.out[b];
                    } else {
                        for (uint32_t b = lane; b < produced; b += 32) dst[b] = sm.out[b];
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) { out_size[c] = produced; status[c] = st; }
    }
}

} // namespace msc
