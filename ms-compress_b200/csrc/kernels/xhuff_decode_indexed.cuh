// xhuff_decode_indexed.cuh -- Xpress-Huffman decoder for streams that come with a restart index (xhuff_index.cuh):
// one WARP per stream, one LANE per 64-symbol segment.
//
// Replaces the reference's xpress_huff_decompress_chunk / HuffmanDecoder / InputBitstream
// (src/xpress_huff_decompress.cpp:39-135, include/mscomp/HuffmanDecoder.h:42-114, include/mscomp/Bitstream.h:34-107) for the
// bulk of the work; the serial walk that finds the segment starts is xhuff_index_kernel (or the encoder, which knows them).
//
// Per stream and per "super-round" of up to 32 consecutive segments of one chunk:
//   * table: the warp builds the chunk's canonical decoder once -- per-length limits, the symbols in canonical order
//     and a 1024-entry first-level table {symbol, length} for codes of up to 10 bits -- in shared memory, shared by all
//     lanes (the lane-per-stream parser needs a private table per lane; that is what limits its occupancy);
//   * parse: lane j starts from entry j's reader state (bit window, valid bits, byte cursor) and decodes its segment
//     8 symbols at a time with exactly the reference's reader arithmetic; every symbol becomes one 32-bit token
//     (offset | length or literal) in a small shared-memory staging area;
//   * apply: the staged tokens are in stream order (segment-major), so the warp executes them 32 at a time: a
//     segmented scan of the lengths gives the positions (every segment's start is known from its lane), literals are
//     stored, matches whose source lies below the first byte of the batch are copied one per lane, the few others (a
//     source inside the batch, a run, a long match) follow in order, copied by the whole warp;
//   * check: a lane must arrive exactly at the next entry's state -- reader and output position -- after exactly the
//     number of symbols its entry announces; so the concatenation of the segments IS the serial parse, whatever the
//     index claims, and an index that does not belong to the stream is reported, never followed into wrong output.
#pragma once
#include "common.cuh"
#include "xhuff_index.cuh"

namespace msc {

constexpr int XHX_Q = 8;                           // symbols per lane per pass
constexpr uint32_t XHX_EMPTY = 0x01000000u;        // token: no symbol (offset 0, "literal" value 256)
constexpr uint32_t XHX_TOKBUF_WORDS = 32 * XHI_SEG_SYMS;       // token scratch of one warp: 32 segments x 64 tokens (8 KB, lives in L2)
constexpr int XHX_WARPS_PER_BLOCK = 4;
constexpr uint32_t XHX_RING = 2048;                // per-warp output window in shared memory (the most recent bytes of the stream)
constexpr uint32_t XHX_RMASK = XHX_RING - 1;
constexpr uint32_t XHX_FLUSH = 512;                // the window is written to global memory in blocks of this size, as soon as they are complete
constexpr uint32_t XHX_SPAN = 512;                // batches that produce more than this take the one-match-at-a-time path
constexpr int XHX_LUT_BITS = 10;
constexpr int XHX_BULK_LEN = 16;
constexpr int32_t ST_INDEX = -1000;                // internal: the index does not describe this stream

struct XhxWarp {
    __align__(16) uint8_t ring[XHX_RING];          // byte at output position x lives at ring[x & XHX_RMASK] while x >= frontier - XHX_RING
    __align__(16) uint8_t slot[32 * 48];           // parse: the next 48 input bytes of every lane (three 16-byte granules from its cursor down-aligned);
                                                   // apply: the first 24 bytes of a lane's slot stage a match source that has left the window
    uint16_t lut[1 << XHX_LUT_BITS];               // first level: symbol | length << 9 (length 0: code longer than 10 bits); bit 15: invalid
    uint16_t syms[XH_SYMS];                        // symbols in canonical (length, symbol) order
    uint32_t lims[16];                             // lims[n], n = 1..15: first 15-bit left-aligned value NOT of length <= n (HuffmanDecoder.h:58-76)
    int32_t rbase[16];                             // rank of a length-n code x: rbase[n] + (x >> (15 - n))
    uint32_t cnt[16];
    uint32_t fill[16];
    uint32_t nused, ok;
};

// tokens: offset in the low half; high half = length - 3 (match) or the byte (literal, offset 0)
__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t* p)
{
#ifndef MSC_EMU
    return __ldcg(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ unsigned long long ldcg_u64(const unsigned long long* p)
{
#ifndef MSC_EMU
    return __ldcg(p);
#else
    return *p;
#endif
}

// HuffmanDecoder::SetCodeLengths (HuffmanDecoder.h:42-91) by the whole warp, from the 256 packed nibbles at `tab`.
__device__ __forceinline__ bool xhx_build_table(XhxWarp& sm, const uint8_t* __restrict__ tab)
{
    const int lane = lane_id();
    if (lane < 16) sm.cnt[lane] = 0;
    __syncwarp();
    for (int i = lane; i < 256; i += 32) {
        const uint32_t b = tab[i];
        atomicAdd(&sm.cnt[b & 15u], 1u);
        atomicAdd(&sm.cnt[b >> 4], 1u);
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t last = 0, pos = 0, prev = 0;
        bool ok = true;
        for (int len = 1; len <= 15; ++len) {
            const uint32_t c = sm.cnt[len];
            pos += prev;                                                 // poss[len] = poss[len-1] + cnts[len-1], cnts[0] counted as 0 (:79-80)
            prev = c;
            sm.fill[len] = pos;
            sm.rbase[len] = (int32_t)pos - (int32_t)(last >> (15 - len));
            if (len < 15) {
                const uint32_t inc = c << (15 - len);
                if (last + inc > 32768u) ok = false;                     // :65, :73 over-subscribed
                last += inc;
                sm.lims[len] = last;
            } else {
                if (last + c > 32768u) ok = false;                       // :76
                sm.lims[15] = 32768u;
                sm.nused = pos + c;
            }
        }
        sm.ok = ok ? 1u : 0u;
    }
    __syncwarp();
    if (!sm.ok) return false;
    // symbols in (length, symbol) order (:83-88): 32 symbols at a time, equal lengths ranked by lane order
    for (int c = 0; c < 16; ++c) {
        const uint32_t s = (uint32_t)(32 * c + lane);
        const uint32_t b = tab[s >> 1];
        const uint32_t len = (s & 1u) ? (b >> 4) : (b & 15u);
        const unsigned m = __match_any_sync(FULL, len);
        const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
        const uint32_t at = len ? sm.fill[len] + rank : 0u;
        if (len) sm.syms[at] = (uint16_t)s;
        __syncwarp();
        if (len && rank == 0) sm.fill[len] += (uint32_t)__popc(m);
        __syncwarp();
    }
    // first level: every 10-bit prefix -> {symbol, length} when the code is at most 10 bits long
    {
        int n = 1;
        for (int i = 0; i < (1 << XHX_LUT_BITS) / 32; ++i) {
            const uint32_t p = (uint32_t)(lane * ((1 << XHX_LUT_BITS) / 32) + i);      // ascending per lane: n only grows
            const uint32_t x = p << (15 - XHX_LUT_BITS);
            while (n < 15 && x >= sm.lims[n]) ++n;
            uint16_t e = 0;
            if (n <= XHX_LUT_BITS) {
                const uint32_t r = (uint32_t)(sm.rbase[n] + (int32_t)(x >> (15 - n)));
                e = r < sm.nused ? (uint16_t)(sm.syms[r] | (n << 9)) : (uint16_t)(0x8000u | (n << 9));
            }
            sm.lut[p] = e;
        }
    }
    __syncwarp();
    return true;
}

// reader of one lane (the reference's InputBitstream state).  Input bytes come from the lane's 48-byte slot in shared memory
// (refilled once per pass of 8 symbols, which without raw length bytes consume at most 30 bytes); anything beyond it -- only
// raw length bytes can get there -- is read from global memory.
struct XhxReader {
    const uint8_t* in;
    uint8_t* slot;             // my slot
    uint32_t slot_ip;          // stream offset of slot[0] (may be "negative": the slot starts at the 16-byte granule below the cursor)
    uint32_t in_len;
    uint32_t mask, bits, ip;
    __device__ __forceinline__ void refill()
    {   // the three 16-byte granules from the cursor's granule on; granules wholly outside the stream's granules are not read
        const uintptr_t g = (uintptr_t)(in + ip) & ~(uintptr_t)15;
        const uintptr_t last = ((uintptr_t)(in + in_len) + 15) & ~(uintptr_t)15;          // end of the last granule holding a stream byte
        slot_ip = ip - (uint32_t)((uintptr_t)(in + ip) - g);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (g + 16u * j < last) v = ld_stream16(reinterpret_cast<const void*>(g + 16u * j));
            *reinterpret_cast<uint4*>(slot + 16 * j) = v;
        }
    }
    __device__ __forceinline__ uint32_t rd8(uint32_t at) const
    {
        const uint32_t d = at - slot_ip;
        return d < 48u ? (uint32_t)slot[d] : (uint32_t)in[at];
    }
    __device__ __forceinline__ uint32_t rd16(uint32_t at) const
    {
        const uint32_t d = at - slot_ip;
        if (d < 47u) return (uint32_t)slot[d] | ((uint32_t)slot[d + 1] << 8);
        return (uint32_t)in[at] | ((uint32_t)in[at + 1] << 8);
    }
    __device__ __forceinline__ bool mask_is_zero() const { return bits == 0 || (mask >> (32 - bits)) == 0; }
    // InputBitstream::Skip (Bitstream.h:59-70) without branches: the 16-bit refill always comes from the slot -- ensure_slot()
    // (one rarely taken branch per symbol) keeps the cursor at least 8 bytes away from the slot's end.
    __device__ __forceinline__ void skip(uint32_t n)
    {
        mask <<= n;                                                      // n <= 15
        bits -= n;
        const bool need = bits < 16u && ip + 2u <= in_len;
        const uint32_t d = ip - slot_ip;
        const uint32_t w = (uint32_t)slot[d] | ((uint32_t)slot[d + 1] << 8);
        if (need) { mask |= w << (16u - bits); bits += 16u; ip += 2u; }
    }
    __device__ __forceinline__ void ensure_slot() { if (ip - slot_ip > 39u) refill(); }
};

// One symbol of a lane's segment -> token; returns false on anything the index walk would not have passed.  Straight-line code
// for literals and ordinary matches alike (a literal is a "match" with no offset bits whose fields are ignored); only codes
// longer than 10 bits and raw length bytes branch.
__device__ __forceinline__ bool xhx_symbol(XhxReader& R, const XhxWarp& sm, uint32_t& tok, uint32_t& len_out)
{
    const uint32_t x = R.mask >> 17;                                     // (bits below the valid ones are zero)
    const uint32_t e = sm.lut[R.mask >> (32 - XHX_LUT_BITS)];
    uint32_t n = (e >> 9) & 15u, sym = e & 511u;
    bool bad = (e & 0x8000u) != 0;
    if (n == 0) {                                                        // 11..15 bits: the canonical limits (HuffmanDecoder.h:105-114)
        n = 11u + (x >= sm.lims[11]) + (x >= sm.lims[12]) + (x >= sm.lims[13]) + (x >= sm.lims[14]);
        const uint32_t r = (uint32_t)(sm.rbase[n] + (int32_t)(x >> (15 - n)));
        bad = r >= sm.nused;
        sym = bad ? 0u : sm.syms[r];
    }
    bad |= n > R.bits;
    n = bad ? 0u : n;
    R.skip(n);
    const bool is_match = sym >= 256u;
    uint32_t len = sym & 15u;
    if (is_match && len == 15u && !bad) {                                // :57-66 length escapes: byte, then u16, then u32
        if (R.in_len - R.ip < 1) return false;
        len = R.rd8(R.ip); R.ip += 1;
        if (len == 0xFFu) {
            if (R.in_len - R.ip < 2) return false;
            len = R.rd16(R.ip); R.ip += 2;
            if (len == 0) {
                if (R.in_len - R.ip < 4) return false;
                len = R.rd16(R.ip) | (R.rd16(R.ip + 2) << 16); R.ip += 4;
            }
            if (len < 15u) return false;
            len -= 15u;
        }
        len += 15u;
        R.ensure_slot();                                                 // (raw bytes moved the cursor)
    }
    uint32_t ob = is_match ? (sym >> 4) & 15u : 0u;
    bad |= ob > R.bits;
    ob = bad ? 0u : ob;
    const uint32_t off = ((R.mask >> 16) >> (16u - ob)) + (1u << ob);
    R.skip(ob);
    bad |= len > 0xFFFFu;                                                // (no room in a token: such a stream takes the index-free decoder)
    tok = is_match ? (off | (len << 16)) : (sym << 16);
    len_out = is_match ? len + 3u : 1u;
    return !bad;
}

// Output side of one stream (all members uniform across the warp).  Every byte is written to the shared-memory window;
// complete 512-byte blocks go to global memory at once (one 16-byte store per lane).  A source byte y is read from the
// window when y >= flushed (invariant: flushed >= frontier - XHX_FLUSH - 32, far inside the window), from global memory
// (where it was stored by this very warp) otherwise.
struct XhxOut {
    uint8_t* o;            // the stream's output in global memory
    uint8_t* ring;
    uint32_t flushed;      // bytes [0, flushed) are in global memory
    bool vec;              // o is 16-byte aligned: blocks leave as 16-byte vectors

    __device__ __forceinline__ void flush_upto(uint32_t frontier)
    {
        const uint32_t lane = (uint32_t)lane_id();
        while (flushed + XHX_FLUSH <= frontier) {
            __syncwarp();
            if (vec) {
                const uint4 v = *reinterpret_cast<const uint4*>(ring + ((flushed + 16u * lane) & XHX_RMASK));
                *reinterpret_cast<uint4*>(o + flushed + 16u * lane) = v;
            } else {
                for (uint32_t i = lane; i < XHX_FLUSH; i += 32) o[flushed + i] = ring[(flushed + i) & XHX_RMASK];
            }
            flushed += XHX_FLUSH;
            __syncwarp();
        }
    }
    __device__ __forceinline__ void flush_tail(uint32_t frontier)
    {
        __syncwarp();
        for (uint32_t i = flushed + (uint32_t)lane_id(); i < frontier; i += 32) o[i] = ring[i & XHX_RMASK];
        flushed = frontier;
        __syncwarp();
    }
    __device__ __forceinline__ uint32_t byte_at(uint32_t y) const
    {
#ifndef MSC_EMU
        return y >= flushed ? (uint32_t)ring[y & XHX_RMASK] : (uint32_t)__ldcg(o + y);
#else
        return y >= flushed ? (uint32_t)ring[y & XHX_RMASK] : (uint32_t)o[y];
#endif
    }
    // Whole-warp copy of one match with byte-serial semantics (arguments uniform); flushes as it goes, so any length works.
    __device__ __forceinline__ void coop_copy(uint32_t pos, uint32_t off, uint32_t len)
    {
        const uint32_t lane = (uint32_t)lane_id();
        uint32_t src = pos - off;
        if (off < 32u && off < len) {
            // A run: the output repeats the `off` bytes before pos.  The first stretch is indexed modulo the period; after that
            // the copy distance can be any multiple of the period: one of at least 32 is used, so 32-byte steps never read
            // what they write.
            const uint32_t off2 = off * ((31u + off) / off);
            const uint32_t head = len < off2 ? len : off2;
            uint32_t v[2];
            for (uint32_t i = lane, k = 0; i < head; i += 32, ++k) v[k] = byte_at(src + i % off);
            for (uint32_t i = lane, k = 0; i < head; i += 32, ++k) ring[(pos + i) & XHX_RMASK] = (uint8_t)v[k];
            __syncwarp();
            pos += head; len -= head;
            src = pos - off2;
            flush_upto(pos);
        }
        for (uint32_t i = 0; i < len; i += 32) {
            if (i + lane < len) ring[(pos + i + lane) & XHX_RMASK] = (uint8_t)byte_at(src + i + lane);
            __syncwarp();
            if (((pos + i + 32u) & (XHX_FLUSH - 1u)) < 32u) flush_upto(pos + (i + 32u < len ? i + 32u : len));
        }
    }
};

// Stream i: in[in_off[i] .. +in_len[i]) -> out[out_off[i] ..], entries index[index_off[i] .. +nseg[i]).
// exp_len / exp_status: what the index walk found (streams with exp_status < 0 are skipped, their status is kept); when
// exp_len is null the index is taken on trust as far as the final length goes (the chain check still holds).
__global__ void __launch_bounds__(32 * XHX_WARPS_PER_BLOCK, 7)
xhuff_decode_indexed_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                            const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                            const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                            const XhSeg* __restrict__ index, const uint64_t* __restrict__ index_off, const uint32_t* __restrict__ nseg,
                            uint32_t* __restrict__ out_len, int32_t* __restrict__ status, int walked, uint32_t* __restrict__ tok_scratch)
{
    __shared__ XhxWarp smem[XHX_WARPS_PER_BLOCK];
    XhxWarp& sm = smem[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * XHX_WARPS_PER_BLOCK;
    uint32_t* const tokbuf = tok_scratch + (size_t)(blockIdx.x * XHX_WARPS_PER_BLOCK + (threadIdx.x >> 5)) * XHX_TOKBUF_WORDS;
    for (uint32_t u = blockIdx.x * XHX_WARPS_PER_BLOCK + (threadIdx.x >> 5); u < count; u += warps_total) {
        if (walked && status[u] < 0) continue;                            // the walk already knows this stream is damaged
        const uint8_t* src = in + in_off[u];
        const uint32_t n_in = in_len[u], cap = out_cap[u];
        uint8_t* o = out + out_off[u];
        const XhSeg* ix = index + index_off[u];
        const uint32_t ns = nseg[u];
        const uint64_t room = index_off[u + 1] - index_off[u];
        int st = ST_OK;
        uint32_t produced = 0;
        XhxOut O;
        O.o = o; O.ring = sm.ring; O.flushed = 0; O.vec = ((uintptr_t)o & 15u) == 0;
        if (ns > room || (!walked && ns == 0 && n_in != 0)) st = ST_INDEX;   // (only the empty stream has no segment)
        uint32_t chunk_start = 0;
        for (uint32_t e0 = 0; e0 < ns && st == ST_OK; ) {
            // ---- the super-round: entries e0 .. e0+cnt-1, all of one chunk -----------------------------------------
            XhSeg E = XhSeg{0, 0, 0, 0};
            if (e0 + lane < ns) E = ix[e0 + lane];
            const unsigned later_chunk = __ballot_sync(FULL, e0 + lane < ns && lane > 0 && (E.info & XHI_NEWCHUNK));
            uint32_t cnt = ns - e0 < 32u ? ns - e0 : 32u;
            if (later_chunk) { const uint32_t c2 = (uint32_t)__ffs((int)later_chunk) - 1u; if (c2 < cnt) cnt = c2; }
            const uint32_t first_info = __shfl_sync(FULL, E.info, 0), first_ip = __shfl_sync(FULL, E.ip, 0);
            if (first_info & XHI_NEWCHUNK) {
                // first segment of a chunk: reader freshly initialised behind the 256-byte table (:143-157, Bitstream.h:44)
                const uint32_t first_pos = __shfl_sync(FULL, E.out_pos, 0), first_mask = __shfl_sync(FULL, E.mask, 0);
                bool ok = first_ip >= 260u && first_ip <= n_in && (first_info & 63u) == 32u && first_pos == produced &&
                          (e0 != 0 || first_ip == 260u);
                if (ok) {
                    const uint8_t* t = src + first_ip - 260u;
                    const uint32_t m0 = ((uint32_t)t[256] << 16) | ((uint32_t)t[257] << 24) | (uint32_t)t[258] | ((uint32_t)t[259] << 8);
                    ok = m0 == first_mask && xhx_build_table(sm, t);
                }
                if (!ok) { st = ST_INDEX; break; }
                chunk_start = first_pos;
            } else if (e0 == 0) { st = ST_INDEX; break; }
            const bool act = (uint32_t)lane < cnt;
            XhxReader R;
            R.in = src; R.in_len = n_in; R.mask = E.mask; R.bits = E.info & 63u; R.ip = E.ip; R.slot = sm.slot + 48 * lane; R.slot_ip = 0x80000000u;
            uint32_t left = act ? (E.info >> 8) & 255u : 0u;             // symbols still to decode
            uint32_t pos = E.out_pos;                                    // my segment's next output position
            uint32_t filled = 0;                                         // passes I wrote tokens for
            bool bad = act && (R.bits > 32u || R.ip > n_in || left == 0 || left > XHI_SEG_SYMS ||
                               (R.bits < 32u && (R.mask << R.bits) != 0u));
            // the first segment continues where the previous super-round ended
            if (lane == 0 && pos != produced) bad = true;
            if (__any_sync(FULL, bad)) { st = ST_INDEX; break; }
            const uint32_t chunk_end = chunk_start + XH_CHUNK;

            // ---- parse: every lane decodes its segment, 8 symbols per pass; the tokens go to this warp's scratch in stream
            //      order (segment-major), 32 bytes per lane and pass --------------------------------------------------------
            for (uint32_t pass = 0; pass < XHI_SEG_SYMS / XHX_Q; ++pass) {
                if (!__any_sync(FULL, left != 0)) break;
                uint32_t t[XHX_Q];
#pragma unroll
                for (int k = 0; k < XHX_Q; ++k) {
                    t[k] = XHX_EMPTY;
                    if (left != 0 && !bad) {
                        // the chunk must not have ended before this symbol (:79), and a literal / match must fit (:87, :113-114)
                        uint32_t len = 0;
                        R.ensure_slot();
                        if (!walked && !(pos < chunk_end || !R.mask_is_zero())) bad = true;     // (the walk has checked this already)
                        else if (!xhx_symbol(R, sm, t[k], len)) bad = true;
                        else {
                            const uint32_t off = t[k] & 0xFFFFu;
                            if (off > pos || cap - pos < len) bad = true;
                            else { pos += len; --left; }
                        }
                        if (bad) t[k] = XHX_EMPTY;
                    }
                }
                uint4* dst = reinterpret_cast<uint4*>(tokbuf + (uint32_t)lane * XHI_SEG_SYMS + pass * XHX_Q);
                dst[0] = make_uint4(t[0], t[1], t[2], t[3]);
                dst[1] = make_uint4(t[4], t[5], t[6], t[7]);
                filled = pass + 1;
            }
            if (__any_sync(FULL, bad)) { st = ST_INDEX; break; }
            __syncwarp();
            // ---- apply: the tokens of the super-round in stream order, 32 at a time (half a segment) --------------------
            const uint32_t nfilled = __reduce_max_sync(FULL, filled) * XHX_Q;            // slots written per segment
            bool slip = false;
            for (uint32_t sg = 0; sg < cnt && !slip; ++sg) {
                const uint32_t seg_pos = __shfl_sync(FULL, E.out_pos, (int)sg);
                if (seg_pos != produced) { slip = true; break; }                          // the segment starts where the one before it ended
                uint32_t tnext = (uint32_t)lane < nfilled ? ldcg_u32(tokbuf + sg * XHI_SEG_SYMS + (uint32_t)lane) : XHX_EMPTY;
                for (uint32_t h = 0; h < nfilled; h += 32) {
                    const uint32_t t = tnext;
                    {   // the next batch's tokens are requested now: this segment's second half, or the next segment's first
                        const uint32_t h2 = h + 32u < nfilled ? h + 32u : 0u, sg2 = h + 32u < nfilled ? sg : sg + 1u;
                        tnext = (sg2 < cnt && h2 + (uint32_t)lane < nfilled) ? ldcg_u32(tokbuf + sg2 * XHI_SEG_SYMS + h2 + (uint32_t)lane) : XHX_EMPTY;
                    }
                    const uint32_t off = t & 0xFFFFu, hi = t >> 16;
                    const uint32_t len = off ? hi + 3u : (hi < 256u ? 1u : 0u);
                    const uint32_t incl = (uint32_t)warp_incl_scan((int)len);
                    const uint32_t p = produced + incl - len;
                    const uint32_t batch_end = produced + __shfl_sync(FULL, incl, 31);
                    // A batch that spans more than XHX_SPAN bytes (long matches) could overwrite window bytes that have not left for
                    // global memory yet: its items are executed one at a time, in order, flushing in between.
                    const bool wide = batch_end - produced > XHX_SPAN;
                    const bool match = off != 0;
                    const uint32_t s0 = p - off;
                    // one per lane: short, not self-overlapping, source wholly below the batch (everything there is final) -- from the
                    // window when it cannot be overwritten during this batch, from global memory otherwise (it left the window long ago)
                    const uint32_t rp = p & XHX_RMASK, rs = s0 & XHX_RMASK;
                    const bool far = (int32_t)s0 < (int32_t)(batch_end - XHX_RING);
                    const bool now = match && len <= (uint32_t)XHX_BULK_LEN && off >= len && s0 + len <= produced && !wide &&
                                     rp + len <= XHX_RING && (far || rs + len <= XHX_RING);       // (a piece that wraps around the window goes the slow way)
                    // far sources: the (up to three) aligned 8-byte words that cover them are requested first and used last
                    unsigned long long w0 = 0, w1 = 0, w2 = 0;
                    const uint32_t fa = (uint32_t)((uintptr_t)(O.o + s0) & 7u);
                    if (now && far) {
                        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(O.o + s0 - fa);
                        w0 = ldcg_u64(w);
                        if (fa + len > 8u) w1 = ldcg_u64(w + 1);
                        if (fa + len > 16u) w2 = ldcg_u64(w + 2);
                    }
                    if (len && !off && !wide) O.ring[rp] = (uint8_t)hi;
                    const uint8_t* q = O.ring + rs;
                    if (now && far) {
                        unsigned long long* f = reinterpret_cast<unsigned long long*>(sm.slot + 48 * lane);
                        f[0] = w0; f[1] = w1; f[2] = w2;
                        q = sm.slot + 48 * lane + fa;
                    }
                    if (now) {
                        // four bytes per step, loads before stores (source and destination never overlap here): the steps' shared-memory
                        // round trips overlap instead of queueing up behind one another
                        uint8_t* d = O.ring + rp;
                        for (uint32_t i = 0; i < len; i += 4) {
                            const uint32_t b0 = q[i], b1 = i + 1 < len ? q[i + 1] : 0u, b2 = i + 2 < len ? q[i + 2] : 0u, b3 = i + 3 < len ? q[i + 3] : 0u;
                            d[i] = (uint8_t)b0;
                            if (i + 1 < len) d[i + 1] = (uint8_t)b1;
                            if (i + 2 < len) d[i + 2] = (uint8_t)b2;
                            if (i + 3 < len) d[i + 3] = (uint8_t)b3;
                        }
                    }
                    __syncwarp();
                    unsigned rest = __ballot_sync(FULL, (match && !now) || (wide && len != 0));
                    while (rest) {
                        const int l = __ffs((int)rest) - 1;
                        rest &= rest - 1;
                        const uint32_t p2 = __shfl_sync(FULL, p, l), f2 = __shfl_sync(FULL, off, l), l2 = __shfl_sync(FULL, len, l), h2 = __shfl_sync(FULL, hi, l);
                        if (f2) O.coop_copy(p2, f2, l2);
                        else { if (lane == 0) O.ring[p2 & XHX_RMASK] = (uint8_t)h2; O.flush_upto(p2 + 1u); }
                    }
                    produced = batch_end;
                    O.flush_upto(produced);
                }
            }
            if (slip) { st = ST_INDEX; break; }
            __syncwarp();
            if (__any_sync(FULL, bad)) { st = ST_INDEX; break; }
            // ---- chain check: every segment ends exactly where the next entry starts -------------------------------
            {
                const uint32_t nxt = e0 + (uint32_t)lane + 1u;
                bool off_chain = false;
                if (act && nxt < ns) {
                    const XhSeg N = ix[nxt];
                    if (N.info & XHI_NEWCHUNK) {
                        // my segment was the last of its chunk: the chunk ends here (:79) and the next table follows the cursor (:127-128)
                        off_chain = N.out_pos != pos || (pos < chunk_end || !R.mask_is_zero()) || N.ip != R.ip + 260u;
                    } else off_chain = N.out_pos != pos || N.ip != R.ip || (N.info & 63u) != R.bits || N.mask != R.mask;
                }
                // the last segment of the stream: what follows must be the end of the stream as the reference sees it --
                // the end-of-stream code with nothing behind it (:83, :129), or a finished chunk at the end of the input (:145-149)
                if (!walked && act && nxt == ns) {
                    XhxReader T = R;
                    T.refill();                           // (the slot has served as match staging since the parse)
                    uint32_t t = 0, l = 0;
                    const bool eos = xhx_symbol(T, sm, t, l) && t == 1u && T.in_len - T.ip == 0 && T.mask_is_zero();
                    const bool chunk_done = !(pos < chunk_end || !R.mask_is_zero());
                    if (!(eos || (chunk_done && R.ip == n_in))) off_chain = true;
                }
                if (__any_sync(FULL, off_chain)) { st = ST_INDEX; break; }
            }
            if (produced != __shfl_sync(FULL, pos, (int)(cnt - 1u))) { st = ST_INDEX; break; }
            e0 += cnt;
        }
        if (st == ST_OK) O.flush_tail(produced);
        if (lane == 0) {
            if (walked) { if (st != ST_OK || produced != out_len[u]) status[u] = ST_INDEX; }
            else { out_len[u] = produced; status[u] = st; }
        }
    }
}

} // namespace msc
