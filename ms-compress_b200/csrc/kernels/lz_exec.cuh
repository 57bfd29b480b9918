// lz_exec.cuh -- small shared pieces of the Xpress / Xpress-Huffman kernels: the LZ77 token word used by the
// encoders' front end and the byte-sequential input window used by the "one lane per stream" decoders.
//
// Both formats are bit/byte-serial per stream, so the parallel axis is the batch: every
// lane of a warp owns one independent buffer.
#pragma once
#include "common.cuh"

namespace msc {

constexpr uint32_t LZQ_MAXLEN = 32768;         // longest single match token the encoders emit

// token: bit31 = match; literal: low 8 bits; match: offset in [0,16), (len-1) in [16,31)
__device__ __forceinline__ uint32_t lzq_lit(uint32_t b) { return b & 0xFFu; }
__device__ __forceinline__ uint32_t lzq_match(uint32_t off, uint32_t len) { return 0x80000000u | ((len - 1) << 16) | off; }

// Byte-sequential reader over one stream in global memory: a 24-byte window held in registers (three aligned 8-byte
// words: two to read from, one already requested for the next slide, so no read ever waits for the load it issues).
// Positions are 32-bit offsets from the start of the stream, so the bookkeeping is single 32-bit instructions;
// reads of <= 4 bytes at the cursor touch memory only when the cursor crosses an 8-byte boundary.
struct ByteWindow {
    const uint8_t* origin;   // 8-byte aligned address at or just below the stream start
    uint32_t skew;           // stream start - origin (0..7)
    uint32_t wb;             // offset (from origin, multiple of 8) of w0
    uint32_t lim;            // first multiple of 8 (from origin) at or beyond the end of the stream
    uint64_t w0, w1, w2;
    __device__ __forceinline__ uint64_t word(uint32_t at8) const { return (at8 < lim) ? *reinterpret_cast<const uint64_t*>(origin + at8) : 0; }
    __device__ __forceinline__ void load(uint32_t at8) { wb = at8; w0 = word(wb); w1 = word(wb + 8); w2 = word(wb + 16); }
    __device__ __forceinline__ void init(const uint8_t* p, const uint8_t* end)
    {
        skew = (uint32_t)((uintptr_t)p & 7u);
        origin = p - skew;
        lim = ((uint32_t)(end - p) + skew + 7u) & ~7u;
        load(0);
    }
    // 4 bytes (little endian) at stream offset `at`; bytes beyond the end of the stream read as 0
    __device__ __forceinline__ uint32_t get32(uint32_t at)
    {
        const uint32_t a = at + skew;
        uint32_t d = a - wb;
        if (d >= 8u) {
            if (d < 16u) {                                       // the usual case: slide by one word
                wb += 8; w0 = w1; w1 = w2;
                w2 = word(wb + 16);
#ifndef MSC_EMU
                // entering a new 128-byte line: ask for the one after the next so that the loads above hit L1 / L2
                if ((wb & 127u) == 0 && wb + 256 < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(origin + wb + 256));
#endif
            } else load(a & ~7u);                                // rewind / long jump (code-length table, new chunk)
            d = a - wb;
        }
        const uint32_t sh = d * 8;                               // 0..56
        uint64_t v = w0 >> sh;
        if (sh > 32) v |= w1 << (64 - sh);
        return (uint32_t)v;
    }
};

} // namespace msc
