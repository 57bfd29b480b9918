// lane_out.cuh -- per-lane LZ77 output cursor for the "one lane per stream" decoders.
//
// Each lane writes its own buffer, so stores cannot be coalesced across lanes; what can be done is to make every
// store wide and to keep the match-source loads off the critical path:
//   * the bytes of the current 8-byte-aligned word are assembled in a register and stored once, when the word is
//     complete; far matches are read 8 bytes at a time (two aligned loads + funnel shift);
//   * the LAST piece (<= 8 bytes) of a far match is only *requested* -- its two raw words wait in registers and are
//     merged into the output when the next item is written, i.e. after the next symbol has been parsed, so the
//     L2 / HBM latency of the load overlaps with that parse (a one-deep software pipeline per lane).
// Invariants: memory is valid for [0, wp) (wp = start of the word under construction); bytes [wp, opos - pk) live in
// `buf`; the last pk bytes of the logical output are the pending piece.
// Replaces the byte-serial copy loops / FAST_COPY of the reference (include/mscomp/internal.h:568-598).
#pragma once
#include "common.cuh"

namespace msc {

struct LaneOut {
    uint8_t* base;          // output buffer of this lane's stream
    uint32_t opos;          // bytes produced (logical: includes the pending piece)
    uint32_t wp;            // bytes [0, wp) are in memory; (base + wp) is 8-byte aligned once past the head
    uint64_t buf;           // bytes [wp, opos - pk), little endian
    uint32_t head;          // number of leading bytes written one by one until (base + x) is 8-byte aligned
    uint64_t pw0, pw1;      // pending piece: raw aligned words of its source
    uint32_t psh, pk;       // byte shift inside pw0, number of pending bytes (0 = none)

    __device__ __forceinline__ void init(uint8_t* b)
    {
        base = b; opos = 0; wp = 0; buf = 0; pw0 = pw1 = 0; psh = 0; pk = 0;
        head = (uint32_t)((8u - ((uintptr_t)b & 7u)) & 7u);
    }
    // append k (1..8) bytes held in the low bytes of v at physical position opos (no pending piece); past the head only
    __device__ __forceinline__ void append(uint64_t v, uint32_t k)
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
    __device__ __forceinline__ void flush_pending()
    {
        if (pk) {
            uint64_t v = pw0 >> (8 * psh);
            if (psh + pk > 8) v |= pw1 << (8 * (8 - psh));
            const uint32_t k = pk;
            pk = 0;
            opos -= k;
            append(v, k);
        }
    }
    __device__ __forceinline__ void put_byte(uint32_t v)
    {
        flush_pending();
        if (opos < head) { base[opos++] = (uint8_t)v; wp = opos; return; }
        buf |= (uint64_t)(v & 0xFFu) << (8 * (opos - wp));
        if (++opos - wp == 8) { *reinterpret_cast<uint64_t*>(base + wp) = buf; wp += 8; buf = 0; }
    }
    // append k (1..8) bytes; falls back to single bytes inside the unaligned head
    __device__ __forceinline__ void put_bytes(uint64_t v, uint32_t k)
    {
        flush_pending();
        if (opos < head) { for (uint32_t i = 0; i < k; ++i) put_byte((uint32_t)(v >> (8 * i))); }
        else append(v, k);
    }
    // byte at absolute output position p < opos (no pending piece)
    __device__ __forceinline__ uint32_t get_byte(uint32_t p) const
    {
        return p >= wp ? (uint32_t)(buf >> (8 * (p - wp))) & 0xFFu : base[p];
    }
    // LZ77 copy of `len` bytes from distance `off` (1 <= off <= opos); byte-serial semantics for overlaps
    __device__ __forceinline__ void copy_match(uint32_t off, uint32_t len)
    {
        flush_pending();
        uint32_t src = opos - off;
        if (opos >= head && src + len <= wp) {
            // far match: the whole source is already in memory
            while (len > 8) {
                const uint8_t* a = base + src;
                const uint32_t sh = (uint32_t)((uintptr_t)a & 7u);
                const uint64_t* w = reinterpret_cast<const uint64_t*>(a - sh);
                uint64_t v = w[0] >> (8 * sh);
                if (sh) v |= w[1] << (8 * (8 - sh));
                append(v, 8);
                src += 8; len -= 8;
            }
            // last piece: request it now, merge it later
            const uint8_t* a = base + src;
            psh = (uint32_t)((uintptr_t)a & 7u);
            const uint64_t* w = reinterpret_cast<const uint64_t*>(a - psh);
            pw0 = w[0];
            if (psh + len > 8) pw1 = w[1];
            pk = len;
            opos += len;
        } else {
            for (; len; --len, ++src) put_byte(get_byte(src));
        }
    }
    __device__ __forceinline__ void emit(bool is_match, uint32_t lit, uint32_t off, uint32_t len)
    {
        if (is_match) copy_match(off, len); else put_byte(lit);
    }
    __device__ __forceinline__ void new_chunk() {}
    __device__ __forceinline__ void before_symbol(uint32_t, uint32_t, uint32_t) {}
    // write out the pending piece and the bytes of the unfinished word
    __device__ __forceinline__ void finish()
    {
        flush_pending();
        for (uint32_t p = wp; p < opos; ++p) base[p] = (uint8_t)(buf >> (8 * (p - wp)));
        wp = opos; buf = 0;
    }
};

} // namespace msc
