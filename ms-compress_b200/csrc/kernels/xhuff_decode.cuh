// xhuff_decode.cuh -- Xpress-Huffman decoder, one lane per independent buffer (32 buffers per warp).
//
// Replaces the reference's xpress_huff_decompress (src/xpress_huff_decompress.cpp:136-162), its chunk loop
// xpress_huff_decompress_chunk (:39-135), the bit reader InputBitstream (include/mscomp/Bitstream.h:34-107) and
// the canonical decoder HuffmanDecoder::SetCodeLengths / DecodeSymbol (include/mscomp/HuffmanDecoder.h:42-103).
//
// Per buffer the format is strictly serial: codes are variable length, raw length bytes are interleaved at the
// reader's BYTE cursor (Bitstream.h:104-106) and a chunk's length is stored nowhere -- so a lane
// keeps the reference's exact reader state (32-bit window, valid-bit count, byte cursor) and walks its buffer
// chunk after chunk; parallelism comes from the batch.  Literals and match copies go through a per-lane output
// cursor that stores whole 8-byte words (lane_out.cuh).  A multi-chunk stream produced by the reference
// (cross-chunk matches, implicit chunk boundaries) decodes on the same path.  The warp re-converges after every
// symbol (the loop condition is a warp vote), so the lanes stay in lock step apart from the literal / match split.
//
// Canonical decoding without a big lookup table: the per-length limits of the left-aligned 15-bit code space
// (`lims`, HuffmanDecoder.h:58-76) live in 14 registers; the code length is 1 + #{k : x >= lims[k]}; the symbol is
// syms[base[n] + (x >> (15-n))].  Canonical order puts the literals of a length before its match symbols, so the
// table stores 8 bits per symbol and a per-length boundary (`litend`) supplies the ninth.
// Shared memory per lane: 512 B symbols + 3 x 16 x 2 B, interleaved across lanes (19 KB per warp, 11 warps per SM).
#pragma once
#include "common.cuh"
#include "lz_exec.cuh"
#include "lane_out.cuh"
#include "lz_apply.cuh"

namespace msc {

constexpr int XH_SYMS = 512;
constexpr int XH_EOS = 256;
constexpr uint32_t XH_CHUNK = 65536;
constexpr int XH_MIN_DATA = 260;

struct XhWarpSmem {
    uint8_t syms[XH_SYMS * 32];          // low 8 bits of the symbol at rank i: syms[i * 32 + lane]
    int16_t base[16 * 32];               // base[n * 32 + lane]; doubles as the per-length counters while building
    uint16_t fill[16 * 32];              // table build only: next free slot per length
    uint16_t litend[16 * 32];            // rank of the first match symbol (>= 256) of each length
};
constexpr int XHD_WARPS_PER_BLOCK = 1;
constexpr size_t XHD_SMEM_BYTES = sizeof(XhWarpSmem) * XHD_WARPS_PER_BLOCK;

struct XhLane {
    // stream
    const uint8_t* in; uint32_t in_len; uint32_t cap;
    // reader state (Bitstream.h:34-44): window, valid bits, byte cursor
    uint32_t mask; int bits; uint32_t ip;
    ByteWindow win;
    // decode tables
    uint32_t lims[15];                   // lims[k] = limit of length k (k = 1..14), left aligned in 15 bits
    uint32_t nused;                      // symbols with a code (ranks >= nused are invalid)
    // progress
    uint32_t chunk_end;
    int status;                          // 0 running, 1 finished OK, <0 error
    bool need_table;
};

__device__ __forceinline__ uint32_t xh_rd16(XhLane& L, uint32_t at) { return L.win.get32(at) & 0xFFFFu; }
__device__ __forceinline__ uint32_t xh_rd32(XhLane& L, uint32_t at) { return L.win.get32(at); }

__device__ __forceinline__ bool xh_mask_is_zero(const XhLane& L) { return L.bits == 0 || (L.mask >> (32 - L.bits)) == 0; }
__device__ __forceinline__ uint32_t xh_peek(const XhLane& L, int n) { return (L.mask >> 16) >> (16 - n); }
__device__ __forceinline__ void xh_skip(XhLane& L, int n)
{   // InputBitstream::Skip (Bitstream.h:59-70)
    L.mask <<= n; L.bits -= n;
    if (L.bits < 16 && L.ip + 2 <= L.in_len) {
        L.mask |= xh_rd16(L, L.ip) << (16 - L.bits);
        L.bits += 16;
        L.ip += 2;
    }
}

// HuffmanDecoder::SetCodeLengths (HuffmanDecoder.h:42-91) from the 256 packed nibbles at in[at..at+256).
__device__ __forceinline__ bool xh_build_table(XhLane& L, XhWarpSmem& sm, int lane, uint32_t at)
{
    for (int l = 0; l < 16; ++l) sm.base[l * 32 + lane] = 0;
    for (int i = 0; i < 256; i += 4) {
        const uint32_t w = xh_rd32(L, at + i);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm.base[((w >> (4 * k)) & 15) * 32 + lane] += 1;
    }
    uint32_t last = 0, pos = 0, cnt_prev = 0;                   // lims[len-1], poss[len], cnts[len-1] (cnts[0] counts as 0)
    bool ok = true;
#pragma unroll
    for (int len = 1; len <= 15; ++len) {
        const uint32_t cnt = (uint32_t)(uint16_t)sm.base[len * 32 + lane];
        pos += cnt_prev;                                        // poss[len] = poss[len-1] + cnts[len-1]   (:79-80)
        cnt_prev = cnt;
        sm.fill[len * 32 + lane] = (uint16_t)pos;
        sm.litend[len * 32 + lane] = (uint16_t)pos;             // corrected below if the length has literals
        sm.base[len * 32 + lane] = (int16_t)((int)pos - (int)(last >> (15 - len)));     // rank = base[n] + (x >> (15-n))
        if (len < 15) {
            const uint32_t inc = cnt << (15 - len);
            if (last + inc > 32768u) ok = false;                // :65, :73 over-subscribed
            last += inc;
            L.lims[len] = last;
        } else {
            if (last + cnt > 32768u) ok = false;                // :76  (lims[15] itself is the whole code space, :61)
            L.nused = pos + cnt;
        }
    }
    if (!ok) return false;
    for (int i = 0; i < 256; i += 4) {                          // symbols in (length, symbol) order (:83-88)
        if (i == 128) {                                         // all literals are placed: the match symbols start here
#pragma unroll
            for (int len = 1; len <= 15; ++len) sm.litend[len * 32 + lane] = sm.fill[len * 32 + lane];
        }
        const uint32_t w = xh_rd32(L, at + i);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t l = (w >> (4 * k)) & 15;
            if (l) {
                const uint32_t f = sm.fill[l * 32 + lane];
                sm.fill[l * 32 + lane] = (uint16_t)(f + 1);
                sm.syms[f * 32 + lane] = (uint8_t)(2 * i + k);
            }
        }
    }
    return true;
}

// DecodeSymbol (HuffmanDecoder.h:93-103): returns the symbol or 0xFFFF
__device__ __forceinline__ uint32_t xh_decode_symbol(XhLane& L, const XhWarpSmem& sm, int lane)
{
    const int r = L.bits;
    const uint32_t x = (r < 15) ? (r == 0 ? 0u : ((L.mask >> (32 - r)) << (15 - r))) : (L.mask >> 17);
    // n = 1 + #{k : x >= lims[k]}: the limits are non-decreasing, so a 4-level binary search over registers finds it
    // with 4 compares and 11 selects (lims[15] is the whole code space: never reached)
    const bool c8 = x >= L.lims[8];
    const bool c4 = x >= (c8 ? L.lims[12] : L.lims[4]);
    const bool c2 = x >= (c8 ? (c4 ? L.lims[14] : L.lims[10]) : (c4 ? L.lims[6] : L.lims[2]));
    const uint32_t l1 = c8 ? (c4 ? (c2 ? 0x8000u : L.lims[13]) : (c2 ? L.lims[11] : L.lims[9]))
                           : (c4 ? (c2 ? L.lims[7] : L.lims[5]) : (c2 ? L.lims[3] : L.lims[1]));
    const int n = 1 + (c8 ? 8 : 0) + (c4 ? 4 : 0) + (c2 ? 2 : 0) + (x >= l1 ? 1 : 0);
    if (n > r) return 0xFFFFu;
    const uint32_t rank = (uint32_t)((int)sm.base[n * 32 + lane] + (int)(x >> (15 - n)));
    xh_skip(L, n);
    if (rank >= L.nused) return 0xFFFFu;                         // unassigned part of an incomplete code (:44, :102)
    return (uint32_t)sm.syms[rank * 32 + lane] | (rank >= (uint32_t)sm.litend[n * 32 + lane] ? 256u : 0u);
}

// One symbol of this lane's stream (xpress_huff_decompress :143-159 around xpress_huff_decompress_chunk :39-135).
template <class Sink>
__device__ __forceinline__ void xh_step(XhLane& L, XhWarpSmem& sm, int lane, Sink& O)
{
    if (L.need_table) {                                         // start of a chunk (:143-157)
        if (L.in_len - L.ip < (uint32_t)XH_MIN_DATA) { L.status = (L.ip == L.in_len) ? 1 : ST_DATA; return; }
        if (!xh_build_table(L, sm, lane, L.ip)) { L.status = ST_DATA; return; }
        L.mask = (xh_rd16(L, L.ip + 256) << 16) | xh_rd16(L, L.ip + 258);
        L.bits = 32;
        L.ip += 260;
        L.chunk_end = O.opos + XH_CHUNK;
        L.need_table = false;
    }
    if (!(O.opos < L.chunk_end || !xh_mask_is_zero(L))) {       // end of chunk (:79, :127-133)
        XhLane T = L;                                            // trial decode of an end-of-stream symbol
        const uint32_t s2 = xh_decode_symbol(T, sm, lane);
        if (s2 == (uint32_t)XH_EOS && T.in_len - T.ip == 0 && xh_mask_is_zero(T)) { L.status = 1; return; }
        L.need_table = true;
        return;
    }
    const uint32_t sym = xh_decode_symbol(L, sm, lane);
    if (sym == 0xFFFFu) { L.status = ST_DATA; return; }
    if (sym == (uint32_t)XH_EOS && L.in_len - L.ip == 0 && xh_mask_is_zero(L)) { L.status = 1; return; }   // :83
    if (sym < 256) {
        if (O.opos == L.cap) { L.status = ST_BUF; return; }
        O.put_byte(sym);
    } else {
        uint32_t len = sym & 15;
        if (len == 15) {
            if (L.in_len - L.ip < 1) { L.status = ST_DATA; return; }
            len = xh_rd16(L, L.ip) & 0xFF; L.ip += 1;
            if (len == 0xFF) {
                if (L.in_len - L.ip < 2) { L.status = ST_DATA; return; }
                len = xh_rd16(L, L.ip); L.ip += 2;
                if (len == 0) {
                    if (L.in_len - L.ip < 4) { L.status = ST_DATA; return; }
                    len = xh_rd32(L, L.ip); L.ip += 4;
                }
                if (len < 15) { L.status = ST_DATA; return; }
                len -= 15;
            }
            len += 15;
        }
        len += 3;
        const int off_bits = (int)((sym >> 4) & 15);
        if (off_bits > L.bits) { L.status = ST_DATA; return; }
        const uint32_t off = xh_peek(L, off_bits) + (1u << off_bits);
        xh_skip(L, off_bits);
        if (O.opos < off) { L.status = ST_DATA; return; }            // :113
        if (L.cap - O.opos < len) { L.status = ST_BUF; return; }     // :114
        O.copy_match(off, len);
    }
}

// Batch decode: buffer i = in[in_off[i] .. +in_len[i]) -> out[out_off[i] .. +out_cap[i]); out_len / status per buffer.
// Sink = LaneOut : single-kernel decoder, the lane writes its own output (no scratch; small batches, huge buffers).
// Sink = LaneTok : stage one of the two-kernel decoder, the lane only appends tokens at tok + 4 * out_off[i];
//                  lz_apply_kernel (lz_apply.cuh) turns them into bytes with one warp per buffer.
template <class Sink>
__device__ __forceinline__ void xhuff_decode_body(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t* __restrict__ tok, uint32_t count,
                                                  const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                                                  const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                                                  uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok)
{
    MSC_DYN_SMEM(smem_raw);
    XhWarpSmem& sm = reinterpret_cast<XhWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * XHD_WARPS_PER_BLOCK;
    for (uint32_t g = blockIdx.x * XHD_WARPS_PER_BLOCK + (threadIdx.x >> 5); g * 32 < count; g += warps_total) {
        const uint32_t i = g * 32 + lane;
        XhLane L;
        L.status = 1; L.in = in; L.in_len = 0; L.cap = 0; L.ip = 0; L.nused = 0;
        L.mask = 0; L.bits = 0; L.chunk_end = 0; L.need_table = true;
#pragma unroll
        for (int k = 0; k < 15; ++k) L.lims[k] = 0x8000u;
        Sink O;
        if (i < count) {
            L.in = in + in_off[i]; L.in_len = in_len[i]; L.cap = out_cap[i];
            sink_init(O, out, tok, out_off[i]);
            L.status = 0;
        } else sink_init(O, out, tok, 0);
        L.win.init(L.in, L.in + L.in_len);
        while (__any_sync(FULL, L.status == 0)) {
            if (L.status == 0) xh_step(L, sm, lane, O);
        }
        O.finish();
        if (i < count) { out_len[i] = O.opos; status[i] = L.status == 1 ? ST_OK : L.status; sink_count(O, ntok, i); }
    }
}

__global__ void __launch_bounds__(32 * XHD_WARPS_PER_BLOCK)
xhuff_decode_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                    const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                    const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                    uint32_t* __restrict__ out_len, int32_t* __restrict__ status)
{
    xhuff_decode_body<LaneOut>(in, out, nullptr, count, in_off, in_len, out_off, out_cap, out_len, status, nullptr);
}

__global__ void __launch_bounds__(32 * XHD_WARPS_PER_BLOCK)
xhuff_parse_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                   const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                   const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                   uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok)
{
    xhuff_decode_body<LaneTok>(in, nullptr, tok, count, in_off, in_len, out_off, out_cap, out_len, status, ntok);
}

} // namespace msc
