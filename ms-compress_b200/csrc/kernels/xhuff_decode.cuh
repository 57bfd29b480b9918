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
// The per-length `base` and `litend` values sit in packed registers and are selected with the predicates of the
// length search, so shared memory holds only the symbols: 512 B per lane, interleaved across lanes (16 KB per warp).
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
                                         // (the first 32 slots of a lane's column hold its counters while it builds a table)
};
// 7 warps x 16 KB per block, 2 blocks per SM: 14 resident warps = 448 buffers per SM, which holds the whole
// 65,536-buffer batch of BASELINE.json configs[3] in ONE wave on 148 SMs (a lane is latency-bound, so waves cost).
constexpr int XHD_WARPS_PER_BLOCK = 7;
constexpr size_t XHD_SMEM_BYTES = sizeof(XhWarpSmem) * XHD_WARPS_PER_BLOCK;

struct XhLane {
    // stream
    const uint8_t* in; uint32_t in_len; uint32_t cap;
    // reader state (Bitstream.h:34-44): window, valid bits, byte cursor
    uint32_t mask; int bits; uint32_t ip;
    ByteWindow win;
    // decode tables, all in registers except the symbols
    uint32_t lims[15];                   // lims[k] = limit of length k (k = 1..14), left aligned in 15 bits
    uint32_t baseP[8];                   // base[n] (16-bit signed) for n-1 = 2k (low half) and 2k+1 (high half)
    uint32_t litendP[8];                 // rank of the first match symbol (>= 256) of each length, packed the same way
    uint32_t nused;                      // symbols with a code (ranks >= nused are invalid)
    // progress
    uint32_t chunk_end;
    int status;                          // 0 running, 1 finished OK, 2 (speculative mode) this chunk is complete, <0 error
    bool need_table;
    // speculative single-chunk mode (xhuff_chunks.cuh): the lane starts at a candidate chunk start with its output position at 0
    bool spec;
    uint32_t reach;                      // how far the chunk's matches reach back before its own first byte (0: self-contained)
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
    // Per-length counters: 16 x 16 bits kept in the first 32 slots of this lane's OWN symbol column (slot 2l = low
    // byte, 2l+1 = high byte), so a lane that rebuilds its table never touches the live tables of the other lanes.
    uint8_t* col = sm.syms + lane;
    for (int f = 0; f < 32; ++f) col[f * 32] = 0;
    uint32_t lit[16];                                           // literal symbols (< 256) per length
    for (int i = 0; i < 256; i += 4) {
        if (i == 128) {
#pragma unroll
            for (int l = 1; l <= 15; ++l) lit[l] = (uint32_t)col[(2 * l) * 32] | ((uint32_t)col[(2 * l + 1) * 32] << 8);
        }
        const uint32_t w = xh_rd32(L, at + i);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t l = (w >> (4 * k)) & 15;
            const uint32_t c = ((uint32_t)col[(2 * l) * 32] | ((uint32_t)col[(2 * l + 1) * 32] << 8)) + 1;
            col[(2 * l) * 32] = (uint8_t)c; col[(2 * l + 1) * 32] = (uint8_t)(c >> 8);
        }
    }
    uint32_t last = 0, pos = 0, cnt_prev = 0;                   // lims[len-1], poss[len], cnts[len-1] (cnts[0] counts as 0)
    uint32_t used_lens = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) { L.baseP[k] = 0; L.litendP[k] = 0; }
#pragma unroll
    for (int len = 1; len <= 15; ++len) {
        const uint32_t c = (uint32_t)col[(2 * len) * 32] | ((uint32_t)col[(2 * len + 1) * 32] << 8);
        pos += cnt_prev;                                        // poss[len] = poss[len-1] + cnts[len-1]   (:79-80)
        cnt_prev = c;
        if (c) used_lens |= 1u << len;
        const uint32_t b16 = (uint32_t)((int)pos - (int)(last >> (15 - len))) & 0xFFFFu;    // rank = base[n] + (x >> (15-n))
        const uint32_t e16 = (pos + lit[len]) & 0xFFFFu;
        const int m = len - 1;
        L.baseP[m >> 1] |= b16 << (16 * (m & 1));
        L.litendP[m >> 1] |= e16 << (16 * (m & 1));
        if (len < 15) {
            const uint32_t inc = c << (15 - len);
            if (last + inc > 32768u) ok = false;                // :65, :73 over-subscribed
            last += inc;
            L.lims[len] = last;
        } else {
            if (last + c > 32768u) ok = false;                  // :76  (lims[15] itself is the whole code space, :61)
            L.nused = pos + c;
        }
    }
    if (!ok) return false;
    // symbols in (length, symbol) order (:83-88): one sweep over the table per code length in use
    uint32_t fill = 0;
    for (int n = 1; n <= 15; ++n) {
        if (!((used_lens >> n) & 1)) continue;
        for (int i = 0; i < 256; i += 4) {
            const uint32_t w = xh_rd32(L, at + i);
            // lanes whose nibble equals n: compare all 8 nibbles at once
            uint32_t eq = w ^ (0x11111111u * (uint32_t)n);      // nibble == 0  <=>  length n
            if (((eq - 0x11111111u) & ~eq & 0x88888888u) == 0) continue;          // no zero nibble in this word
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (((eq >> (4 * k)) & 15) == 0) { sm.syms[fill * 32 + lane] = (uint8_t)(2 * i + k); ++fill; }
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
    // with 4 compares and a few selects (lims[15] is the whole code space: never reached)
    const bool c8 = x >= L.lims[8];
    const bool c4 = x >= (c8 ? L.lims[12] : L.lims[4]);
    const bool c2 = x >= (c8 ? (c4 ? L.lims[14] : L.lims[10]) : (c4 ? L.lims[6] : L.lims[2]));
    const uint32_t l1 = c8 ? (c4 ? (c2 ? 0x8000u : L.lims[13]) : (c2 ? L.lims[11] : L.lims[9]))
                           : (c4 ? (c2 ? L.lims[7] : L.lims[5]) : (c2 ? L.lims[3] : L.lims[1]));
    const bool c1 = x >= l1;
    const int n = 1 + (c8 ? 8 : 0) + (c4 ? 4 : 0) + (c2 ? 2 : 0) + (c1 ? 1 : 0);
    if (n > r) return 0xFFFFu;
    // the same predicates select base[n] and litend[n] out of their packed registers
    const uint32_t bp = c8 ? (c4 ? (c2 ? L.baseP[7] : L.baseP[6]) : (c2 ? L.baseP[5] : L.baseP[4]))
                           : (c4 ? (c2 ? L.baseP[3] : L.baseP[2]) : (c2 ? L.baseP[1] : L.baseP[0]));
    const uint32_t ep = c8 ? (c4 ? (c2 ? L.litendP[7] : L.litendP[6]) : (c2 ? L.litendP[5] : L.litendP[4]))
                           : (c4 ? (c2 ? L.litendP[3] : L.litendP[2]) : (c2 ? L.litendP[1] : L.litendP[0]));
    const int base = (int)(int16_t)(uint16_t)(c1 ? (bp >> 16) : bp);
    const uint32_t litend = c1 ? (ep >> 16) : (ep & 0xFFFFu);
    const uint32_t rank = (uint32_t)(base + (int)(x >> (15 - n)));
    xh_skip(L, n);
    if (rank >= L.nused) return 0xFFFFu;                         // unassigned part of an incomplete code (:44, :102)
    return (uint32_t)sm.syms[rank * 32 + lane] | (rank >= litend ? 256u : 0u);
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
        O.new_chunk();
    }
    if (!(O.opos < L.chunk_end || !xh_mask_is_zero(L))) {       // end of chunk (:79, :127-133)
        XhLane T = L;                                            // trial decode of an end-of-stream symbol
        const uint32_t s2 = xh_decode_symbol(T, sm, lane);
        if (s2 == (uint32_t)XH_EOS && T.in_len - T.ip == 0 && xh_mask_is_zero(T)) { L.status = 1; return; }
        if (L.spec) { L.status = 2; return; }                    // the next chunk (table at L.ip) is another lane's
        L.need_table = true;
        return;
    }
    O.before_symbol(L.mask, (uint32_t)L.bits, L.ip);
    const uint32_t sym = xh_decode_symbol(L, sm, lane);
    if (sym == 0xFFFFu) { L.status = ST_DATA; return; }
    if (sym == (uint32_t)XH_EOS && L.in_len - L.ip == 0 && xh_mask_is_zero(L)) { L.status = 1; return; }   // :83
    // Literals and matches share one instruction stream (a literal is handled as a "match" with 0 offset bits whose
    // length/offset are ignored), so the lanes of the warp only part ways for the rare long-length escape codes.
    const bool is_match = sym >= 256;
    uint32_t len = sym & 15;
    if (is_match && len == 15) {                                // :57-66 length escapes: byte, then u16, then u32
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
    len = is_match ? len + 3 : 1u;
    const int off_bits = is_match ? (int)((sym >> 4) & 15) : 0;
    if (off_bits > L.bits) { L.status = ST_DATA; return; }
    const uint32_t off = xh_peek(L, off_bits) + (1u << off_bits);
    if (off_bits) xh_skip(L, off_bits);
    if (is_match && O.opos < off) {                                      // :113
        if (!L.spec) { L.status = ST_DATA; return; }
        L.reach = max(L.reach, off - O.opos);                            // (whether that much output precedes the chunk is the stitcher's call)
    }
    if (L.cap - O.opos < len) { L.status = ST_BUF; return; }             // :114 (and :87 for a literal)
    O.emit(is_match, sym, off, len);
}

// Batch decode: buffer i = in[in_off[i] .. +in_len[i]) -> out[out_off[i] .. +out_cap[i]); out_len / status per buffer.
// Sink = LaneOut : single-kernel decoder, the lane writes its own output (no scratch; small batches, huge buffers).
// Sink = LaneTok : stage one of the two-kernel decoder, the lane only appends tokens at tok + 4 * out_off[i];
//                  lz_apply_kernel (lz_apply.cuh) turns them into bytes with one warp per buffer.
template <class Sink>
__device__ __forceinline__ void xhuff_decode_body(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t* __restrict__ tok, uint32_t count,
                                                  const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                                                  const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                                                  uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok,
                                                  int only_status = 0, unsigned long long* __restrict__ handled = nullptr,
                                                  uint32_t* __restrict__ end_ip = nullptr, uint32_t* __restrict__ reach = nullptr)
{
    MSC_DYN_SMEM(smem_raw);
    XhWarpSmem& sm = reinterpret_cast<XhWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * XHD_WARPS_PER_BLOCK;
    for (uint32_t g = blockIdx.x * XHD_WARPS_PER_BLOCK + (threadIdx.x >> 5); g * 32 < count; g += warps_total) {
        const uint32_t i = g * 32 + lane;
        // only_status != 0: second-chance pass -- only the streams another decoder left marked with that status are decoded
        const bool mine = i < count && (only_status == 0 || status[i] == only_status);
        if (only_status != 0 && !__any_sync(FULL, mine)) continue;
        if (mine && handled) atomicAdd(handled, 1ull);
        XhLane L;
        L.status = 1; L.in = in; L.in_len = 0; L.cap = 0; L.ip = 0; L.nused = 0;
        L.mask = 0; L.bits = 0; L.chunk_end = 0; L.need_table = true;
        L.spec = end_ip != nullptr; L.reach = 0;
#pragma unroll
        for (int k = 0; k < 15; ++k) L.lims[k] = 0x8000u;
#pragma unroll
        for (int k = 0; k < 8; ++k) { L.baseP[k] = 0; L.litendP[k] = 0; }
        Sink O;
        if (mine) {
            L.in = in + in_off[i]; L.in_len = in_len[i]; L.cap = out_cap[i];
            sink_init(O, out, tok, out_off[i]);
            L.status = 0;
        } else sink_init(O, out, tok, 0);
        L.win.init(L.in, L.in + L.in_len);
        while (__any_sync(FULL, L.status == 0)) {
            if (L.status == 0) xh_step(L, sm, lane, O);
        }
        O.finish();
        if (mine) {
            out_len[i] = O.opos; sink_count(O, ntok, i);
            if (end_ip) { status[i] = L.status == 1 ? 0 : (L.status == 2 ? 1 : L.status); end_ip[i] = L.ip; reach[i] = L.reach; }   // 0 stream ends here, 1 chunk complete
            else status[i] = L.status == 1 ? ST_OK : L.status;
        }
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

// The same decoder as a second chance: only streams whose status equals only_status are decoded (and counted).
__global__ void __launch_bounds__(32 * XHD_WARPS_PER_BLOCK)
xhuff_decode_marked_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                           const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                           const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                           uint32_t* __restrict__ out_len, int32_t* __restrict__ status, int only_status, unsigned long long* __restrict__ handled)
{
    xhuff_decode_body<LaneOut>(in, out, nullptr, count, in_off, in_len, out_off, out_cap, out_len, status, nullptr, only_status, handled);
}

__global__ void __launch_bounds__(32 * XHD_WARPS_PER_BLOCK)
xhuff_parse_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                   const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                   const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                   uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok)
{
    xhuff_decode_body<LaneTok>(in, nullptr, tok, count, in_off, in_len, out_off, out_cap, out_len, status, ntok);
}

// Speculative mode: buffer i starts at a candidate chunk start; the lane parses ONE chunk into tokens (xhuff_chunks.cuh).
// status[i]: 0 the stream ends inside / right after this chunk, 1 the chunk is complete and another table follows at end_ip[i]
// (relative to the buffer), < 0 not a decodable chunk.
__global__ void __launch_bounds__(32 * XHD_WARPS_PER_BLOCK)
xhuff_parse_chunks_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                          const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                          const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                          uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok,
                          uint32_t* __restrict__ end_ip, uint32_t* __restrict__ reach)
{
    xhuff_decode_body<LaneTok>(in, nullptr, tok, count, in_off, in_len, out_off, out_cap, out_len, status, ntok, 0, nullptr, end_ip, reach);
}

} // namespace msc
