// lznt1_host.h -- host-side logic of the one-shot lznt1_decompress that does not depend on where the chunks are decoded:
// the serial header walk (small streams; large ones get their index from the GPU, kernels/lznt1_index.cuh) and the replay of
// the reference's sequential driver over per-chunk results, which yields its status code.  Shared by the product
// (mscomp_shim.cu) and by the CPU test host (tests/emu/stream_host.cpp: same logic over the oracle's chunk decoder).
#pragma once
#include "../../include/mscomp.h"
#include <stdint.h>
#include <vector>

namespace msc_host {

enum Lznt1Tail { T_CLEAN, T_ZERO, T_GARBAGE, T_TERM_OK, T_TERM_BAD, T_TRUNC, T_BADSIG };

// Header walk: the format stores no chunk index and each header gives only the distance to the next one
// (src/lznt1_decompress.cpp:127-135), so an unindexed stream is a linked list.  offs = header offsets of the whole chunks,
// *ip = bytes they cover, return value = how the walk ended.
inline Lznt1Tail lznt1_walk(const uint8_t* in, size_t in_len, std::vector<uint64_t>& offs, size_t* ip_out)
{
    Lznt1Tail tail = T_CLEAN;
    size_t ip = 0;
    bool stopped = false;
    while (in_len - ip >= 2) {
        const unsigned h = (unsigned)in[ip] | ((unsigned)in[ip + 1] << 8);
        if (h == 0) { tail = (in_len - ip == 2) ? T_TERM_OK : T_TERM_BAD; stopped = true; break; }
        const size_t sz = (h & 0x0FFF) + 3;
        if (sz > in_len - ip) { tail = T_TRUNC; stopped = true; break; }
        if ((h & 0x7000) != 0x3000) { tail = T_BADSIG; stopped = true; break; }
        offs.push_back(ip);
        ip += sz;
    }
    if (!stopped) tail = (in_len - ip == 0) ? T_CLEAN : (in[ip] == 0 ? T_ZERO : T_GARBAGE);
    *ip_out = ip;
    return tail;
}

// Replays the reference's sequential driver (src/lznt1_decompress.cpp:254-273 + internal.h:616-630) over the
// per-chunk results to obtain its status code.  size_of / status_of give chunk c's decoded size and status.
template <class FS, class FT>
MSCompStatus lznt1_replay(size_t nchunks, size_t cap, Lznt1Tail tail, FS size_of, FT status_of, size_t* total)
{
    size_t op = 0;
    for (size_t c = 0; c < nchunks; ++c) {
        if (cap - op == 0) return MSCOMP_BUF_ERROR;        // output full while input remains
        if (status_of(c) < 0) return MSCOMP_DATA_ERROR;
        const size_t sz = size_of(c);
        if (sz > cap - op) return MSCOMP_BUF_ERROR;        // decoded bytes left over in the bounce buffer
        op += sz;
    }
    *total = op;
    switch (tail) {
    case T_CLEAN: case T_ZERO: return MSCOMP_OK;
    case T_GARBAGE: return MSCOMP_BUF_ERROR;
    default: break;
    }
    if (cap - op == 0) return MSCOMP_BUF_ERROR;            // the loop never looks at the next header (:254)
    switch (tail) {
    case T_TERM_OK: return MSCOMP_OK;
    case T_TRUNC: return MSCOMP_BUF_ERROR;
    default: return MSCOMP_DATA_ERROR;                     // T_TERM_BAD, T_BADSIG
    }
}

} // namespace msc_host
