// lane_out.cuh -- per-lane LZ77 output cursor for the "one lane per stream" decoders.
//
// Each lane writes its own buffer, so stores cannot be coalesced across lanes; what can be done is to make every
// store wide.  The cursor keeps the bytes of the current 8-byte-aligned word in a register and stores the word
// once, when it is complete; far matches are read 8 bytes at a time (two aligned loads + funnel shift).  Memory is
// guaranteed valid for [0, wp) (wp = start of the word under construction); bytes [wp, opos) live in `buf`.
// Replaces the byte-serial copy loops / FAST_COPY of the reference (include/mscomp/internal.h:568-598).
#pragma once
#include "common.cuh"

namespace msc {

struct LaneOut {
    uint8_t* base;          // output buffer of this lane's stream
    uint32_t opos;          // bytes produced
    uint32_t wp;            // bytes [0, wp) are in memory; (base + wp) is 8-byte aligned once past the head
    uint64_t buf;           // bytes [wp, opos), little endian
    uint32_t head;          // number of leading bytes written one by one until (base + x) is 8-byte aligned

    __device__ __forceinline__ void init(uint8_t* b)
    {
        base = b; opos = 0; wp = 0; buf = 0;
        head = (uint32_t)((8u - ((uintptr_t)b & 7u)) & 7u);
    }
    __device__ __forceinline__ void put_byte(uint32_t v)
    {
        if (opos < head) { base[opos++] = (uint8_t)v; wp = opos; return; }
        buf |= (uint64_t)(v & 0xFFu) << (8 * (opos - wp));
        if (++opos - wp == 8) { *reinterpret_cast<uint64_t*>(base + wp) = buf; wp += 8; buf = 0; }
    }
    // append k (1..8) bytes held in the low bytes of v; only valid past the head
    __device__ __forceinline__ void put_bytes(uint64_t v, uint32_t k)
    {
        const uint32_t fill = opos - wp;
        if (k < 8) v &= (1ull << (8 * k)) - 1;
        buf |= v << (8 * fill);
        opos += k;
        if (fill + k >= 8) {
            *reinterpret_cast<uint64_t*>(base + wp) = buf;
            wp += 8;
            buf = fill ? (v >> (8 * (8 - fill))) : 0;
        }
    }
    // byte at absolute output position p < opos
    __device__ __forceinline__ uint32_t get_byte(uint32_t p) const
    {
        return p >= wp ? (uint32_t)(buf >> (8 * (p - wp))) & 0xFFu : base[p];
    }
    // 8 bytes starting at absolute output position p, all of them inside [0, wp)
    __device__ __forceinline__ uint64_t get8_mem(uint32_t p, uint32_t need) const
    {
        const uint8_t* a = base + p;
        const uint32_t sh = (uint32_t)((uintptr_t)a & 7u);
        const uint64_t* w = reinterpret_cast<const uint64_t*>(a - sh);
        uint64_t v = w[0] >> (8 * sh);
        if (sh + need > 8) v |= w[1] << (8 * (8 - sh));
        return v;
    }
    // LZ77 copy of `len` bytes from distance `off` (1 <= off <= opos); byte-serial semantics for overlaps
    __device__ __forceinline__ void copy_match(uint32_t off, uint32_t len)
    {
        uint32_t src = opos - off;
        if (opos >= head && src + len <= wp) {
            // far match: the whole source is already in memory
            while (len) {
                const uint32_t k = len < 8 ? len : 8;
                put_bytes(get8_mem(src, k), k);
                src += k; len -= k;
            }
        } else {
            for (; len; --len, ++src) put_byte(get_byte(src));
        }
    }
    // write out the bytes of the unfinished word
    __device__ __forceinline__ void finish()
    {
        for (uint32_t p = wp; p < opos; ++p) base[p] = (uint8_t)(buf >> (8 * (p - wp)));
        wp = opos; buf = 0;
    }
};

} // namespace msc
