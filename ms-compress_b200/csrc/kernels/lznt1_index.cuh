// lznt1_index.cuh -- chunk index of an LZNT1 stream, built on the GPU.
//
// The format stores no chunk index: a header gives only the distance to the next header, so an unindexed stream is a
// linked list.  The reference walks it serially, one header per chunk (src/lznt1_decompress.cpp:254-273 driver,
// :127-135 header read); on the host that walk costs one cache miss per chunk (~120 ns -> ~30 ms per GiB of output) and
// would dominate lznt1_decompress.  Here the list is ranked in parallel instead:
//   1  tables   the stream is cut into segments of S bytes.  A header can start at most 4097 bytes past a segment
//               boundary (a chunk is at most 0xFFF + 3 = 4098 bytes), so for each segment and EVERY possible entry offset
//               0..4097 a thread walks the segment (staged in shared memory) and records where that walk enters the next
//               segment -- or END if it terminates inside (terminator, bad signature, truncated chunk, end of stream).
//   2  compose  groups of 32 consecutive segment tables are composed into one table per group.
//   3  entries  one thread chains the group tables from the known entry of the first segment (a few hundred dependent
//               loads), then every group expands its 32 segment entries in parallel.
//   4  count    one thread per segment walks the TRUE chain through its segment and counts its chunks; exclusive scan
//               -> first chunk number of every segment; the thread that reaches the end of the chain records how it
//               ended (the reference's status depends on it).
//   5  write    the same walk again, writing chunk offsets.
// Steps work on a range of segments and carry a small state from range to range, so the index of a stream can be built
// while the stream is still being uploaded.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int LZI_THREADS = 1024;
constexpr uint32_t LZI_CAND = 4098;        // entry offsets 0..4097
constexpr uint32_t LZI_STRIDE = 4104;      // table row pitch (entries)
constexpr uint32_t LZI_END = 0xFFFFu;      // the walk terminates inside the segment
constexpr uint32_t LZI_GROUP = 32;
constexpr uint32_t LZI_SEG_BYTES = 65536;  // default segment size (any multiple of 16 >= 4112 works)

// how the chain ended; same meaning and order as Lznt1Tail in mscomp_shim.cu
enum { LZI_CLEAN = 0, LZI_ZERO, LZI_GARBAGE, LZI_TERM_OK, LZI_TERM_BAD, LZI_TRUNC, LZI_BADSIG };

struct Lznt1IndexState {
    uint32_t entry;        // offset of the first header inside the next segment (valid while !done)
    uint32_t done;         // the chain has ended
    uint32_t tail;         // LZI_* (valid when done)
    uint32_t next_entry;   // scratch: entry after the range being processed
    uint64_t ip;           // bytes consumed by whole chunks (valid when done)
    uint64_t nchunks;      // chunks indexed so far
    uint64_t range_chunks; // chunks in the range processed last
};

// One header.  Returns -1 for a chunk (sz = its size), LZI_TERM_OK.. for the reference's stop conditions, or -2 when fewer
// than two bytes remain (end of stream).  Order of the checks = the host walk it replaces / the reference's driver.
template <class RD>
__device__ __forceinline__ int lzi_header(uint64_t in_len, uint64_t abs, RD rd16, uint32_t& sz)
{
    if (abs + 2 > in_len) return -2;
    const uint32_t h = rd16();
    if (h == 0) return (in_len - abs == 2) ? LZI_TERM_OK : LZI_TERM_BAD;
    sz = (h & 0x0FFFu) + 3u;
    if ((uint64_t)sz > in_len - abs) return LZI_TRUNC;
    if ((h & 0x7000u) != 0x3000u) return LZI_BADSIG;
    return -1;
}

// ---- 1: per-segment tables -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LZI_THREADS)
lznt1_index_tables_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t seg0, uint32_t nseg, uint32_t S,
                          uint16_t* __restrict__ T1)
{
    MSC_DYN_SMEM(sm);                                   // S + 16 bytes
    const int tid = (int)threadIdx.x;
    for (uint32_t s = blockIdx.x; s < nseg; s += gridDim.x) {
        const uint64_t segbase = (uint64_t)(seg0 + s) * S;
        const bool aligned = (((uintptr_t)(in + segbase)) & 15u) == 0;
        for (uint32_t i = (uint32_t)tid * 16; i < S + 16; i += LZI_THREADS * 16) {
            if (aligned && segbase + i + 16 <= in_len) {
                *reinterpret_cast<uint4*>(sm + i) = *reinterpret_cast<const uint4*>(in + segbase + i);
            } else {
                for (uint32_t k = 0; k < 16; ++k) sm[i + k] = (segbase + i + k < in_len) ? in[segbase + i + k] : (uint8_t)0;
            }
        }
        __syncthreads();
        for (uint32_t e = (uint32_t)tid; e < LZI_CAND; e += LZI_THREADS) {
            uint32_t p = e, res;
            for (;;) {
                if (p >= S) { res = p - S; break; }
                uint32_t sz = 0;
                const int r = lzi_header(in_len, segbase + p, [&]() { return (uint32_t)sm[p] | ((uint32_t)sm[p + 1] << 8); }, sz);
                if (r != -1) { res = LZI_END; break; }
                p += sz;
            }
            T1[(uint64_t)s * LZI_STRIDE + e] = (uint16_t)res;
        }
        __syncthreads();
    }
}

// ---- 2: one table per group of LZI_GROUP segments ------------------------------------------------------------------------
__global__ void __launch_bounds__(LZI_THREADS)
lznt1_index_compose_kernel(const uint16_t* __restrict__ T1, uint32_t nseg, uint16_t* __restrict__ T2)
{
    const uint32_t g = blockIdx.x;
    const uint32_t s0 = g * LZI_GROUP, s1 = min(nseg, s0 + LZI_GROUP);
    for (uint32_t e = threadIdx.x; e < LZI_CAND; e += LZI_THREADS) {
        uint32_t v = e;
        for (uint32_t s = s0; s < s1 && v != LZI_END; ++s) v = T1[(uint64_t)s * LZI_STRIDE + v];
        T2[(uint64_t)g * LZI_STRIDE + e] = (uint16_t)v;
    }
}

// ---- 3: the true entry of every segment (single block) -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
lznt1_index_entries_kernel(const uint16_t* __restrict__ T1, const uint16_t* __restrict__ T2, uint32_t nseg,
                           Lznt1IndexState* __restrict__ st, uint16_t* __restrict__ seg_entry, uint16_t* __restrict__ grp_entry)
{
    const uint32_t ng = (nseg + LZI_GROUP - 1) / LZI_GROUP;
    if (threadIdx.x == 0) {
        uint32_t v = st->done ? LZI_END : st->entry;
        for (uint32_t g = 0; g < ng; ++g) {
            grp_entry[g] = (uint16_t)v;
            if (v != LZI_END) v = T2[(uint64_t)g * LZI_STRIDE + v];
        }
        st->next_entry = v;
    }
    __syncthreads();
    for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) {
        uint32_t v = grp_entry[g];
        const uint32_t s1 = min(nseg, (g + 1) * LZI_GROUP);
        for (uint32_t s = g * LZI_GROUP; s < s1; ++s) {
            seg_entry[s] = (uint16_t)v;
            if (v != LZI_END) v = T1[(uint64_t)s * LZI_STRIDE + v];
        }
    }
}

// Walks the true chain through segment (seg0 + s); f(abs) is called for every chunk.  Returns the number of chunks;
// end_kind >= 0 (with end_ip) when the chain ends inside this segment.
template <class F>
__device__ __forceinline__ uint32_t lzi_walk_segment(const uint8_t* __restrict__ in, uint64_t in_len, uint64_t segbase, uint32_t S,
                                                     uint32_t entry, F f, int& end_kind, uint64_t& end_ip)
{
    end_kind = -1;
    if (entry == LZI_END) return 0;
    uint64_t p = segbase + entry;
    const uint64_t seg_end = segbase + S;
    uint32_t n = 0;
    int r = -1;
    while (p < seg_end) {
        uint32_t sz = 0;
        r = lzi_header(in_len, p, [&]() { return (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8); }, sz);
        if (r != -1) break;
        f(p);
        ++n;
        p += sz;
    }
    if (r == -1 && seg_end >= in_len) r = -2;            // last segment of the stream, left exactly at its end
    if (r == -2) { end_kind = (in_len - p == 0) ? LZI_CLEAN : (in[p] == 0 ? LZI_ZERO : LZI_GARBAGE); end_ip = p; }
    else if (r >= 0) { end_kind = r; end_ip = p; }
    return n;
}

// ---- 4: chunks per segment, first chunk number of every segment, end of the chain (single block) -----------------------------
__global__ void __launch_bounds__(LZI_THREADS)
lznt1_index_count_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t seg0, uint32_t nseg, uint32_t S,
                         const uint16_t* __restrict__ seg_entry, uint32_t* __restrict__ seg_base, Lznt1IndexState* __restrict__ st,
                         Lznt1IndexState* mirror)   // optional copy of the state in mapped host memory (saves a device-to-host memcpy
{                                                   // that would queue behind the bulk downloads on the copy engine)
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < nseg; t0 += LZI_THREADS) {
        const uint32_t s = t0 + (uint32_t)tid;
        uint32_t n = 0;
        if (s < nseg) {
            int kind; uint64_t ip = 0;
            n = lzi_walk_segment(in, in_len, (uint64_t)(seg0 + s) * S, S, seg_entry[s], [](uint64_t) {}, kind, ip);
            if (kind >= 0) { st->done = 1; st->tail = (uint32_t)kind; st->ip = ip; }     // at most one thread of all ranges
        }
        const uint32_t incl = (uint32_t)warp_incl_scan((int)n);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t woff = carry;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        if (s < nseg) seg_base[s] = woff + incl - n;
        __syncthreads();
        if (tid == LZI_THREADS - 1) carry = woff + incl;
        __syncthreads();
    }
    if (tid == 0) {
        st->range_chunks = carry;
        st->nchunks += carry;
        st->entry = st->next_entry;
        if (mirror) { *mirror = *st; __threadfence_system(); }
    }
}

// ---- 5: chunk offsets --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lznt1_index_write_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t seg0, uint32_t nseg, uint32_t S,
                         const uint16_t* __restrict__ seg_entry, const uint32_t* __restrict__ seg_base, uint64_t* __restrict__ offs)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += gridDim.x * blockDim.x) {
        uint64_t* o = offs + seg_base[s];
        int kind; uint64_t ip;
        lzi_walk_segment(in, in_len, (uint64_t)(seg0 + s) * S, S, seg_entry[s], [&](uint64_t abs) { *o++ = abs; }, kind, ip);
    }
}

} // namespace msc
