// xhuff_chunk_warp.cuh -- ONE Xpress-Huffman chunk parsed into tokens as fast as one chain of dependent instructions goes.
//
// The one-shot xpress_huff_decompress finds its parallelism between the chunks of a stream (xhuff_chunks.cuh); inside a chunk the
// ≈ 12,000 symbols are one serial chain, and with few chunks the call's time IS that chain.  The batch parser
// (xhuff_decode.cuh, one lane per stream, speculative mode) is built for 65,536 lanes at once: no lookup table (512 bytes of shared
// memory per lane is all it may use), the code length from a binary search over registers -- ≈ 1,300 cycles per symbol when a
// lane runs alone.  Here a WARP owns the chunk: all lanes build the canonical decoder of xhuff_decode_indexed.cuh (1,024-entry
// first-level table, HuffmanDecoder.h:42-91) in shared memory, then lane 0 walks the symbols with the indexed decoder's
// straight-line symbol routine (xhx_symbol: one table load per code of up to 10 bits, input from a 48-byte shared-memory slot).
// Same contract as xhuff_parse_chunks_kernel (the reference's chunk loop src/xpress_huff_decompress.cpp:39-135 in speculative
// mode): buffer i starts at a candidate chunk start; tokens (lz_apply.cuh format) go to tok + out_off[i]; status[i] = 0 the stream
// ends inside / right after this chunk, 1 the chunk is complete and another table follows at end_ip[i], < 0 not a decodable chunk;
// reach[i] = how far the chunk's matches reach back before its own first byte.
#pragma once
#include "common.cuh"
#include "lz_apply.cuh"
#include "xhuff_decode.cuh"
#include "xhuff_decode_indexed.cuh"

namespace msc {

constexpr int XHW_WARPS_PER_BLOCK = 2;

__global__ void __launch_bounds__(32 * XHW_WARPS_PER_BLOCK)
xhuff_parse_chunks_warp_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                               const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                               const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                               uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok,
                               uint32_t* __restrict__ end_ip, uint32_t* __restrict__ reach)
{
    __shared__ XhxWarp smem[XHW_WARPS_PER_BLOCK];
    XhxWarp& sm = smem[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * XHW_WARPS_PER_BLOCK;
    for (uint32_t i = blockIdx.x * XHW_WARPS_PER_BLOCK + (threadIdx.x >> 5); i < count; i += warps_total) {
        const uint8_t* s = in + in_off[i];
        const uint32_t n = in_len[i], cap = out_cap[i];
        if (n == 0 && cap == 0) continue;                                         // (a padding slot of the launch layout)
        uint32_t* t = tok + out_off[i];
        int st = 0;
        if (n < (uint32_t)XH_MIN_DATA) st = ST_DATA;                              // :143-157 (a candidate always has a table)
        else if (!xhx_build_table(sm, s)) st = ST_DATA;
        __syncwarp();
        uint32_t opos = 0, nt = 0, rch = 0, ip_end = 0;
        if (st == 0 && lane == 0) {
            XhxReader R;
            R.in = s; R.slot = sm.slot; R.in_len = n;
            R.ip = 256; R.refill();
            R.mask = (R.rd16(256) << 16) | R.rd16(258);
            R.bits = 32; R.ip = 260;
            R.ensure_slot();
            for (;;) {
                uint32_t tk, ln;
                if (opos >= XH_CHUNK && R.mask_is_zero()) {                       // end of chunk (:79, :127-133)
                    XhxReader T = R;                                              // trial decode of an end-of-stream symbol
                    const bool ok = xhx_symbol(T, sm, tk, ln);
                    st = (ok && tk == 1u && ln == 3u && T.in_len - T.ip == 0 && T.mask_is_zero()) ? 1 : 2;
                    break;
                }
                if (!xhx_symbol(R, sm, tk, ln)) { st = ST_DATA; break; }
                // symbol 256 (offset 1, length 3 as a match) is the end-of-stream mark when nothing follows it (:83)
                if (tk == 1u && ln == 3u && R.in_len - R.ip == 0 && R.mask_is_zero()) { st = 1; break; }
                const uint32_t off = tk & 0xFFFFu;                                // 0: a literal
                if (off && opos < off) rch = max(rch, off - opos);                // :113 (whether that much output precedes the chunk is the stitcher's call)
                if (cap - opos < ln) { st = ST_BUF; break; }                      // :114, :87
                if (off) {
                    uint32_t left = ln;
                    while (left) { const uint32_t l = left > LZQ_MAXLEN ? LZQ_MAXLEN : left; t[nt++] = lzt_match(off, l); left -= l; }
                } else t[nt++] = lzt_lit(tk >> 16);
                opos += ln;
                R.ensure_slot();
            }
            ip_end = R.ip;
            out_len[i] = opos; ntok[i] = nt; end_ip[i] = ip_end; reach[i] = rch;
            status[i] = st == 1 ? 0 : (st == 2 ? 1 : st);
        } else if (lane == 0) {
            out_len[i] = 0; ntok[i] = 0; end_ip[i] = 0; reach[i] = 0; status[i] = st;
        }
        __syncwarp();
    }
}

} // namespace msc
