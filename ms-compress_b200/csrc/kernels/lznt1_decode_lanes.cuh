// lznt1_decode_lanes.cuh -- LZNT1 chunk decoder, one LANE per 4 KiB chunk (32 chunks per warp).
//
// Same contract as lznt1_decode_kernel (lznt1_decode.cuh) -- replaces src/lznt1_decompress.cpp:37-121 and the per-chunk
// part of :122-209 -- but with the opposite mapping: a lane walks its own chunk exactly like the reference's checked
// loop (:77-116), so there is no chase / scan / watermark machinery and almost every instruction does useful work for
// 32 chunks at once.  Input bytes come through a 16-byte register window (8-byte loads), output through LaneOut
// (8-byte word stores, far matches read 8 bytes at a time from the chunk's own output in L2).  Runs of literal items
// (consecutive 0 flags) are moved as one unit.  Used for large batches, where chunks outnumber the lanes of the GPU;
// the warp-per-chunk kernel serves small batches.
#pragma once
#include "common.cuh"
#include "lz_exec.cuh"
#include "lane_out.cuh"
#include "lznt1_decode.cuh"

namespace msc {

constexpr int LZL_WARPS_PER_BLOCK = 4;

__global__ void __launch_bounds__(32 * LZL_WARPS_PER_BLOCK)
lznt1_decode_lanes_kernel(const uint8_t* __restrict__ in, uint64_t in_len, const uint64_t* __restrict__ chunk_off, uint32_t nchunks,
                          uint8_t* __restrict__ out, uint32_t* __restrict__ out_size, int32_t* __restrict__ status)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x;
    // all lanes of a warp run the same number of rounds (c - lane is warp-uniform), the tail lanes just idle
    for (uint32_t c = first; c - (threadIdx.x & 31) < nchunks; c += stride) {
        int st = ST_OK;
        LaneOut O;
        O.init(out + (uint64_t)(c < nchunks ? c : 0) * LZNT1_CHUNK);
        ByteWindow win;
        const uint8_t* src = in;
        uint32_t n = 0, ip = 0, f = 1u;
        bool active = false, raw = false;
        if (c < nchunks) {
            const uint64_t start = chunk_off[c];
            if (start + 2 > in_len) st = ST_BUF;
            else {
                const uint32_t hdr = (uint32_t)in[start] | ((uint32_t)in[start + 1] << 8);
                n = (hdr & 0x0FFF) + 1;                                  // payload bytes
                if (hdr == 0) st = ST_DATA;
                else if (start + 2 + n > in_len) st = ST_BUF;            // truncated chunk (reference: internal.h:627)
                else if ((hdr & 0x7000) != 0x3000) st = ST_DATA;         // src/lznt1_decompress.cpp:152
                else { src = in + start + 2; active = true; raw = !(hdr & 0x8000); }
            }
        }
        win.init(src, src + (active ? n : 0));
        // One step per iteration and lane: a flag byte fetch plus one item (token or run of literals), or 8 stored bytes.
        // The loop condition is a warp vote, so the lanes re-converge after every step.
        while (__any_sync(FULL, active)) {
            if (!active) continue;
            if (raw) {                                                   // stored chunk :188-206
                const uint32_t k = min(n - ip, 8u);
                uint64_t v = win.get32(ip);
                if (k > 4) v |= (uint64_t)win.get32(ip + 4) << 32;
                if (O.opos < O.head) { for (uint32_t i = 0; i < k; ++i) O.put_byte((uint32_t)(v >> (8 * i))); }
                else O.put_bytes(v, k);
                ip += k;
                if (ip >= n) active = false;
                continue;
            }
            if (f == 1u) {                                               // next flag byte (:80-82)
                if (ip >= n) { active = false; continue; }
                f = (win.get32(ip) & 0xFFu) | 0x100u;              // flag bits, LSB first, with a sentinel
                ++ip;
            }
            if (ip == n) { active = false; continue; }                   // :86 input ends inside the group
            if (f & 1u) {                                                // token
                if (ip + 2 > n) { st = ST_DATA; active = false; continue; }          // :90
                const uint32_t t = win.get32(ip) & 0xFFFFu;
                ip += 2;
                int S, lim;
                lznt1_shift_limit((int)O.opos, S, lim);                  // :91
                const uint32_t len = (t & ((1u << S) - 1)) + 3, off = (t >> S) + 1;
                if (off > O.opos || O.opos + len > (uint32_t)LZNT1_CHUNK) { st = ST_DATA; active = false; continue; }   // :98-99
                O.copy_match(off, len);
                f >>= 1;
            } else {                                                     // run of literal items
                uint32_t k = (uint32_t)__ffs((int)f) - 1;                // consecutive 0 flags (<= 8: the sentinel stops it)
                k = min(k, n - ip);
                if (O.opos + k > (uint32_t)LZNT1_CHUNK) { st = ST_DATA; active = false; continue; }
                uint64_t v = win.get32(ip);
                if (k > 4) v |= (uint64_t)win.get32(ip + 4) << 32;
                if (O.opos < O.head) { for (uint32_t i = 0; i < k; ++i) O.put_byte((uint32_t)(v >> (8 * i))); }
                else O.put_bytes(v, k);
                ip += k;
                f >>= k;
            }
        }
        O.finish();
        if (c < nchunks) { out_size[c] = st == ST_OK ? O.opos : 0; status[c] = st; }
    }
}

} // namespace msc
