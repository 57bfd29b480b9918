// xhuff_chunk_warp.cuh -- ONE Xpress-Huffman chunk parsed into tokens by a whole warp.
//
// The one-shot xpress_huff_decompress finds its parallelism between the chunks of a stream (xhuff_chunks.cuh); inside a chunk the
// ≈ 12,000 symbols are one serial chain, and with few chunks the call's time IS that chain.  The batch parser
// (xhuff_decode.cuh, one lane per stream, speculative mode) is built for 65,536 lanes at once: no lookup table (512 bytes of shared
// memory per lane is all it may use), the code length from a binary search over registers -- ≈ 1,300 cycles per symbol when a
// lane runs alone.  Here a WARP owns the chunk: all lanes build the canonical decoder of xhuff_decode_indexed.cuh (1,024-entry
// first-level table, HuffmanDecoder.h:42-91) in shared memory, and then
//   * PAR = false: lane 0 walks the symbols with the indexed decoder's straight-line symbol routine (xhx_symbol: one table load
//     per code of up to 10 bits, input from a 48-byte shared-memory slot) -- ≈ 500 cycles per symbol;
//   * PAR = true (the default): the warp decodes by speculation on the start bit.  The reader's state between two symbols is
//     (window, valid bits, cursor) (Bitstream.h:34-44), and as long as no raw length byte is read the bits to come are simply the
//     window followed by the 16-bit words at the cursor.  So lane j decodes the symbol that WOULD start j bits ahead (code, offset
//     bits -- up to 30 bits of the look-ahead every lane holds), the true chain 0 -> adv(0) -> ... is followed with one shuffle
//     per symbol (which also carries the length: the output positions add up along the way), its lanes write their tokens, and
//     the reader state after the last one follows
//     from the number of bits consumed: a refill happens whenever fewer than 16 bits are left (Bitstream.h:59-70), so with b = valid
//     bits - consumed, k = the refills that bring b back into [16, 31]: valid bits = b + 16 k, cursor += 2 k.  ≈ 4 symbols per round.
//     Anything that is not plain -- a length escape (raw bytes move the cursor), an invalid code, the end of the chunk or of the
//     room, the last 16 bytes of the input -- stops the chain in front of it and is done by lane 0 with the serial routine.
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
#ifdef MSC_EMU
inline unsigned long long g_xhw_round_syms = 0, g_xhw_serial_syms = 0;            // emulator diagnostics: symbols taken by speculative rounds / one at a time
#endif

// progress of one chunk (uniform across the warp in the PAR = true kernel)
struct XhwState {
    uint32_t opos, nt, rch;
    int st;                  // 0 running, 1 the stream ends here, 2 the chunk is complete, < 0 error
};

// One symbol by the serial routine (the loop body of xh_step, speculative mode); the caller is one lane.
__device__ __forceinline__ void xhw_serial_step(XhxReader& R, const XhxWarp& sm, uint32_t cap, uint32_t* __restrict__ t, XhwState& S)
{
    uint32_t tk, ln;
    if (S.opos >= XH_CHUNK && R.mask_is_zero()) {                                 // end of chunk (:79, :127-133)
        XhxReader T = R;                                                          // trial decode of an end-of-stream symbol
        const bool ok = xhx_symbol(T, sm, tk, ln);
        S.st = (ok && tk == 1u && ln == 3u && T.in_len - T.ip == 0 && T.mask_is_zero()) ? 1 : 2;
        return;
    }
    if (!xhx_symbol(R, sm, tk, ln)) { S.st = ST_DATA; return; }
    // symbol 256 (offset 1, length 3 as a match) is the end-of-stream mark when nothing follows it (:83)
    if (tk == 1u && ln == 3u && R.in_len - R.ip == 0 && R.mask_is_zero()) { S.st = 1; return; }
    const uint32_t off = tk & 0xFFFFu;                                            // 0: a literal
    if (off && S.opos < off) S.rch = max(S.rch, off - S.opos);                    // :113 (whether that much output precedes the chunk is the stitcher's call)
    if (cap - S.opos < ln) { S.st = ST_BUF; return; }                             // :114, :87
    if (off) {
        uint32_t left = ln;
        while (left) { const uint32_t l = left > LZQ_MAXLEN ? LZQ_MAXLEN : left; t[S.nt++] = lzt_match(off, l); left -= l; }
    } else t[S.nt++] = lzt_lit(tk >> 16);
    S.opos += ln;
    R.ensure_slot();
}

template <bool PAR>
__device__ __forceinline__ void xhuff_parse_chunks_warp_body(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
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
        XhwState S;
        S.opos = 0; S.nt = 0; S.rch = 0; S.st = 0;
        if (n < (uint32_t)XH_MIN_DATA) S.st = ST_DATA;                            // :143-157 (a candidate always has a table)
        else if (!xhx_build_table(sm, s)) S.st = ST_DATA;
        __syncwarp();
        uint32_t ip_end = 0;
        if (!PAR) {
            if (S.st == 0 && lane == 0) {
                XhxReader R;
                R.in = s; R.slot = sm.slot; R.in_len = n;
                R.ip = 256; R.refill();
                R.mask = (R.rd16(256) << 16) | R.rd16(258);
                R.bits = 32; R.ip = 260;
                R.ensure_slot();
                while (S.st == 0) xhw_serial_step(R, sm, cap, t, S);
                ip_end = R.ip;
            }
        } else if (S.st == 0) {
            // the reader, uniform across the warp; the slot is shared, lane 0 refills it
            XhxReader R;
            R.in = s; R.slot = sm.slot; R.in_len = n; R.slot_ip = 0;
            R.ip = 256;
            if (lane == 0) R.refill();
            R.slot_ip = (uint32_t)__shfl_sync(FULL, R.slot_ip, 0);
            __syncwarp();
            R.mask = (R.rd16(256) << 16) | R.rd16(258);
            R.bits = 32; R.ip = 260;
            while (S.st == 0) {
                if (R.ip - R.slot_ip > 39u) {                                    // ensure_slot, by lane 0 for everybody
                    __syncwarp();
                    if (lane == 0) R.refill();
                    R.slot_ip = (uint32_t)__shfl_sync(FULL, R.slot_ip, 0);
                    __syncwarp();
                }
                uint32_t taken = 0;                                              // symbols of this round's chain
                if (R.ip + 16u <= n && R.bits >= 16u && S.opos < XH_CHUNK) {
                    // ---- a speculative round.  The bits to come: the window's valid bits, then the four words at the cursor (P).
                    // V0 = bits [0, 64) of that string, V1 = bits [32, 96): a lane needs at most 30 bits from an offset below 32,
                    // the reader's next window at most 31 bits from an offset below 62. ----
                    const uint32_t d = R.ip - R.slot_ip;
                    unsigned long long P = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        P |= (unsigned long long)((uint32_t)R.slot[d + 2 * k] | ((uint32_t)R.slot[d + 2 * k + 1] << 8)) << (48 - 16 * k);
                    const unsigned long long V0 = ((unsigned long long)R.mask << 32) | (P >> R.bits);                 // (bits in 16..32)
                    const unsigned long long V1 = (V0 << 32) | (unsigned long long)(uint32_t)((P << (64u - R.bits)) >> 32);
                    // lane j: the symbol that would start j bits ahead
                    const unsigned long long Vj = V0 << lane;
                    const uint32_t x = (uint32_t)(Vj >> 49);                     // 15 bits
                    const uint32_t e = sm.lut[x >> (15 - XHX_LUT_BITS)];
                    uint32_t nb = (e >> 9) & 15u, sym = e & 511u;
                    bool stop = (e & 0x8000u) != 0;
                    if (nb == 0) {                                               // 11..15 bits: the canonical limits (HuffmanDecoder.h:105-114)
                        nb = 11u + (x >= sm.lims[11]) + (x >= sm.lims[12]) + (x >= sm.lims[13]) + (x >= sm.lims[14]);
                        const uint32_t r = (uint32_t)(sm.rbase[nb] + (int32_t)(x >> (15 - nb)));
                        stop = r >= sm.nused;
                        sym = stop ? 0u : sm.syms[r];
                    }
                    const bool is_match = sym >= 256u;
                    stop |= is_match && (sym & 15u) == 15u;                      // a length escape: raw bytes at the cursor, lane 0's business
                    const uint32_t ob = is_match ? (sym >> 4) & 15u : 0u;
                    const uint32_t obits = (uint32_t)((Vj << nb) >> 49) >> (15u - ob);              // the ob bits behind the code
                    const uint32_t off = is_match ? obits + (1u << ob) : 0u;
                    const uint32_t len = is_match ? (sym & 15u) + 3u : 1u;
                    // the true chain from bit 0, one shuffle per symbol: adv | len << 8 | stop << 16
                    const uint32_t mine3 = (nb + ob) | (len << 8) | (stop ? 0x10000u : 0u);
                    uint32_t M = 0, cur = 0, run = S.opos, before = 0;
                    while (cur < 32u) {
                        const uint32_t v = __shfl_sync(FULL, mine3, (int)cur);
                        const uint32_t l = (v >> 8) & 0xFFu;
                        // an escape or an invalid code; the chunk's end; the room's end: the chain stops in front of it
                        if ((v >> 16) || run >= XH_CHUNK || cap - run < l) break;
                        if ((uint32_t)lane == cur) before = run;
                        M |= 1u << cur;
                        run += l;
                        cur += v & 0xFFu;
                    }
                    const uint32_t used = cur;                                    // bits consumed
                    const bool on = (M >> lane) & 1u;
                    taken = (uint32_t)__popc(M);
                    if (taken) {
                        const uint32_t rank = (uint32_t)__popc(M & ((1u << lane) - 1u));
                        if (on) t[S.nt + rank] = is_match ? lzt_match(off, len) : lzt_lit(sym);
                        const uint32_t r = (on && is_match && before < off) ? off - before : 0u;
                        S.rch = max(S.rch, __reduce_max_sync(FULL, r));
                        S.opos = run;
                        S.nt += taken;
#ifdef MSC_EMU
                        if (lane == 0) g_xhw_round_syms += taken;
#endif
                        // the reader after `used` bits (Bitstream.h:59-70: a refill whenever fewer than 16 bits are left)
                        const int b0 = (int)R.bits - (int)used;
                        const uint32_t k = b0 >= 16 ? 0u : (uint32_t)(31 - b0) >> 4;
                        const uint32_t b = (uint32_t)(b0 + 16 * (int)k);          // 16..31 (32 only for a round that consumed nothing)
                        R.ip += 2u * k;
                        R.bits = b;
                        const unsigned long long Vn = used < 32u ? V0 << used : V1 << (used - 32u);
                        R.mask = (uint32_t)(Vn >> 32) & (0xFFFFFFFFu << (32u - b));
                    }
                }
                if (!taken) {                                                    // one symbol the serial way
                    __syncwarp();
                    if (lane == 0) xhw_serial_step(R, sm, cap, t, S);
#ifdef MSC_EMU
                    if (lane == 0) g_xhw_serial_syms += 1;
#endif
                    R.mask = __shfl_sync(FULL, R.mask, 0); R.bits = __shfl_sync(FULL, R.bits, 0); R.ip = __shfl_sync(FULL, R.ip, 0);
                    R.slot_ip = __shfl_sync(FULL, R.slot_ip, 0);
                    S.opos = __shfl_sync(FULL, S.opos, 0); S.nt = __shfl_sync(FULL, S.nt, 0); S.rch = __shfl_sync(FULL, S.rch, 0);
                    S.st = __shfl_sync(FULL, S.st, 0);
                    __syncwarp();
                }
            }
            ip_end = R.ip;
        }
        if (lane == 0) {
            out_len[i] = S.opos; ntok[i] = S.nt; end_ip[i] = ip_end; reach[i] = S.rch;
            status[i] = S.st == 1 ? 0 : (S.st == 2 ? 1 : S.st);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * XHW_WARPS_PER_BLOCK)
xhuff_parse_chunks_warp_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                               const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                               const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                               uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok,
                               uint32_t* __restrict__ end_ip, uint32_t* __restrict__ reach)
{
    xhuff_parse_chunks_warp_body<true>(in, tok, count, in_off, in_len, out_off, out_cap, out_len, status, ntok, end_ip, reach);
}

// The same with lane 0 walking every symbol (MSCOMP_XHUFF_CHUNK_PARSE=serial): the measured step before the speculative rounds.
__global__ void __launch_bounds__(32 * XHW_WARPS_PER_BLOCK)
xhuff_parse_chunks_warp_serial_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                               const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                               const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                               uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok,
                               uint32_t* __restrict__ end_ip, uint32_t* __restrict__ reach)
{
    xhuff_parse_chunks_warp_body<false>(in, tok, count, in_off, in_len, out_off, out_cap, out_len, status, ntok, end_ip, reach);
}

} // namespace msc
