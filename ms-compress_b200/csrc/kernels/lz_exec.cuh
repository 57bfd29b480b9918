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

// Byte-sequential reader over global memory with a 16-byte window held in registers (two aligned 8-byte words).
// Reads of <= 4 bytes at the cursor never touch memory except when the cursor crosses an 8-byte boundary.
struct ByteWindow {
    const uint8_t* base;     // aligned-down address of w0
    uint64_t w0, w1;
    const uint8_t* lim;      // first 8-byte aligned address at or beyond the end of the buffer
    __device__ __forceinline__ void init(const uint8_t* p, const uint8_t* end)
    {
        base = p - ((uintptr_t)p & 7u);
        lim = end + ((8u - ((uintptr_t)end & 7u)) & 7u);
        w0 = (base < lim) ? *reinterpret_cast<const uint64_t*>(base) : 0;
        w1 = (base + 8 < lim) ? *reinterpret_cast<const uint64_t*>(base + 8) : 0;
    }
    // make [p, p+4) readable from the window; cheap when p moves forward by a few bytes
    __device__ __forceinline__ void seek(const uint8_t* p)
    {
        if (p >= base && p < base + 8) return;
        if (p >= base + 8 && p < base + 16) {                      // the usual case: slide by one word
            base += 8; w0 = w1;
            w1 = (base + 8 < lim) ? *reinterpret_cast<const uint64_t*>(base + 8) : 0;
#ifndef MSC_EMU
            // entering a new 128-byte line: ask for the one after the next so that the loads above hit L1 / L2
            if ((((uintptr_t)base) & 127u) == 0 && base + 256 < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + 256));
#endif
            return;
        }
        init(p, lim);                                               // rewind / long jump (code-length table, new chunk)
    }
    __device__ __forceinline__ uint32_t get32(const uint8_t* p)
    {   // 4 bytes at p (little endian), p in [base, base+8)
        const uint32_t sh = (uint32_t)(p - base) * 8;            // 0..56
        uint64_t v = w0 >> sh;
        if (sh > 32) v |= w1 << (64 - sh);
        return (uint32_t)v;
    }
};

} // namespace msc
