// xpress_decode.cuh -- plain Xpress (LZ77) decoder, one lane per independent buffer (32 buffers per warp).
//
// Replaces the reference's xpress_decompress (src/xpress_decompress.cpp:405-462: checked loop :425-460; the
// unchecked fast path INFLATE_FAST :119-198 computes the same thing) and its symbol reader (:62-113).
//
// The stream is one unbroken sequence of 32-item flag words interleaved with item bytes, plus a shared length
// nibble whose partner can be arbitrarily far away: it has to be parsed serially, so a lane
// walks its own buffer; its output goes through a cursor that stores whole 8-byte words (lane_out.cuh).
#pragma once
#include "common.cuh"
#include "lz_exec.cuh"
#include "lane_out.cuh"
#include "lz_apply.cuh"
#include "xpress_encode.cuh"   // XpressSeg

namespace msc {

constexpr int XPD_WARPS_PER_BLOCK = 4;

struct XpLane {
    const uint8_t* in; uint32_t in_len; uint32_t cap;
    ByteWindow win;
    uint32_t ip;
    uint32_t flags; bool flagged, have_flags;
    uint32_t half_val; bool has_half;        // value of the byte whose high nibble is still unused (:88-92)
    uint32_t stop_ip;                        // indexed decode: this lane's segment ends when ip reaches it (else 0xFFFFFFFF)
    int status;                              // 0 running, 1 finished OK, < 0 error
};

__device__ __forceinline__ uint32_t xp_rd32(XpLane& L, uint32_t at) { return L.win.get32(at); }

// all 1 bits above all 0 bits (src/xpress_decompress.cpp:41)
__device__ __forceinline__ bool xp_set_bits_are_highest(uint32_t x) { x = ~x; return ((x + 1) & x) == 0; }

// One item of this lane's stream.
template <class Sink>
__device__ __forceinline__ void xp_step(XpLane& L, Sink& O)
{
    do {
        if (L.ip >= L.stop_ip) { L.status = (L.ip == L.stop_ip) ? 1 : ST_DATA; break; }     // end of this lane's segment
        if (!L.have_flags) {                                            // start a fragment (:425-430)
            if (L.ip + 4 > L.in_len) { L.status = ST_DATA; break; }     // :461
            const uint32_t f = xp_rd32(L, L.ip);
            L.flagged = (f & 0x80000000u) != 0;
            L.flags = (f << 1) | 1;
            L.ip += 4;
            L.have_flags = true;
        }
        if (L.ip == L.in_len) {                                         // :433-438 end of stream
            L.status = (L.flagged && xp_set_bits_are_highest(L.flags)) ? 1 : ST_DATA;
            break;
        }
        if (L.flagged) {                                                // _READ_SYMBOL :85-112
            if (L.ip + 2 > L.in_len) { L.status = ST_DATA; break; }
            const uint32_t sym = xp_rd32(L, L.ip) & 0xFFFFu;
            L.ip += 2;
            const uint32_t off = (sym >> 3) + 1;
            uint32_t len = sym & 7;
            if (len == 7) {
                if (L.has_half) { len = L.half_val >> 4; L.has_half = false; }
                else if (L.ip == L.in_len) { L.status = ST_DATA; break; }
                else { L.half_val = xp_rd32(L, L.ip) & 0xFF; L.has_half = true; len = L.half_val & 0xF; L.ip += 1; }
                if (len == 0xF) {
                    if (L.ip == L.in_len) { L.status = ST_DATA; break; }
                    len = xp_rd32(L, L.ip) & 0xFF; L.ip += 1;
                    if (len == 0xFF) {
                        if (L.ip + 2 > L.in_len) { L.status = ST_DATA; break; }
                        len = xp_rd32(L, L.ip) & 0xFFFF; L.ip += 2;
                        if (len == 0) {
                            if (L.ip + 4 > L.in_len) { L.status = ST_DATA; break; }
                            len = xp_rd32(L, L.ip); L.ip += 4;
                        }
                        if (len < 0xF + 0x7) { L.status = ST_DATA; break; }
                        len -= 0xF + 0x7;
                    }
                    len += 0xF;
                }
                len += 0x7;
            }
            len += 0x3;
            if (off > O.opos) { L.status = ST_DATA; break; }            // :442
            if ((uint64_t)O.opos + len > L.cap) { L.status = ST_BUF; break; }     // :443
            O.copy_match(off, len);
        } else {
            if (O.opos == L.cap) { L.status = ST_BUF; break; }          // :455
            O.put_byte(xp_rd32(L, L.ip));
            L.ip += 1;
        }
        L.flagged = (L.flags & 0x80000000u) != 0;
        L.flags <<= 1;
        if (L.flags == 0) L.have_flags = false;
    } while (false);
}

// Sink = LaneOut: single-kernel decoder; Sink = LaneTok: stage one of the two-kernel decoder (see xhuff_decode.cuh).
template <class Sink>
__device__ __forceinline__ void xpress_decode_body(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t* __restrict__ tok, uint32_t count,
                                                   const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                                                   const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                                                   uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok)
{
    const int lane = lane_id();
    const int wib = (int)(threadIdx.x >> 5);
    const uint32_t warps_total = gridDim.x * XPD_WARPS_PER_BLOCK;
    for (uint32_t g = blockIdx.x * XPD_WARPS_PER_BLOCK + wib; g * 32 < count; g += warps_total) {
        const uint32_t i = g * 32 + lane;
        XpLane L;
        L.in = in; L.in_len = 0; L.cap = 0; L.ip = 0; L.flags = 0; L.flagged = false; L.have_flags = false;
        L.half_val = 0; L.has_half = false; L.status = 1; L.stop_ip = 0xFFFFFFFFu;
        Sink O;
        if (i < count) {
            L.in = in + in_off[i]; L.in_len = in_len[i]; L.cap = out_cap[i];
            sink_init(O, out, tok, out_off[i]);
            L.status = 0;
            if (L.in_len < 5) {                                         // :414-418 (sic)
                uint32_t w = 0;
                if (L.in_len == 4) w = (uint32_t)L.in[0] | ((uint32_t)L.in[1] << 8) | ((uint32_t)L.in[2] << 16) | ((uint32_t)L.in[3] << 24);
                L.status = (L.in_len == 0 || (L.in_len == 4 && w != 0xFFFFFFFFu)) ? 1 : ST_DATA;
            }
        } else sink_init(O, out, tok, 0);
        L.win.init(L.in, L.in + L.in_len);
        while (__any_sync(FULL, L.status == 0)) {      // warp vote: the lanes re-converge after every item
            if (L.status == 0) xp_step(L, O);
        }
        O.finish();
        if (i < count) { out_len[i] = O.opos; status[i] = L.status == 1 ? ST_OK : L.status; sink_count(O, ntok, i); }
    }
}

__global__ void __launch_bounds__(32 * XPD_WARPS_PER_BLOCK)
xpress_decode_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                     const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                     const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                     uint32_t* __restrict__ out_len, int32_t* __restrict__ status)
{
    xpress_decode_body<LaneOut>(in, out, nullptr, count, in_off, in_len, out_off, out_cap, out_len, status, nullptr);
}

__global__ void __launch_bounds__(32 * XPD_WARPS_PER_BLOCK)
xpress_parse_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                    const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                    const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                    uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok)
{
    xpress_decode_body<LaneTok>(in, nullptr, tok, count, in_off, in_len, out_off, out_cap, out_len, status, ntok);
}

// ---- one WARP per stream ------------------------------------------------------------------------------------------------
// A stream without a restart index (any reference-made stream) is a serial chain, but not a byte-serial one: the 32 items
// behind a flag word are 1 (literal) or 2 (match symbol) bytes each, plus the rare length extensions.  A warp parses a whole
// flag group at a time: lane i takes item i, its offset is the prefix sum of the base sizes, and the extended matches of
// the group (length field 7) are resolved in order -- the lowest unresolved one always has an exact offset, it reads its
// extension bytes, and the lanes behind it shift.  ~150 warp instructions per 32 items instead of ~115 per item and lane:
// the latency of ONE stream drops by the warp width, which is what the one-shot xpress_decompress of a single buffer and
// batches of few, large reference-made streams need.  Same statuses as xp_step (the first failing item in order decides).
__device__ __forceinline__ uint32_t xpw_rd(const uint8_t* __restrict__ in, uint32_t at, uint32_t nbytes)
{
    uint32_t v = 0;
    for (uint32_t k = 0; k < nbytes; ++k) v |= (uint32_t)__ldg(in + at + k) << (8 * k);
    return v;
}

__global__ void __launch_bounds__(32 * XPD_WARPS_PER_BLOCK)
xpress_parse_warp_kernel(const uint8_t* __restrict__ in_base, uint32_t* __restrict__ tok_base, uint32_t count,
                         const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len_a,
                         const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                         uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint32_t* __restrict__ ntok_out)
{
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * XPD_WARPS_PER_BLOCK;
    for (uint32_t u = blockIdx.x * XPD_WARPS_PER_BLOCK + (threadIdx.x >> 5); u < count; u += warps_total) {
        const uint8_t* in = in_base + in_off[u];
        const uint32_t n = in_len_a[u], cap = out_cap[u];
        uint32_t* tok = tok_base + out_off[u];
        uint32_t ip = 0, opos = 0, ntok = 0, half_val = 0;
        bool has_half = false;
        int st = 0;                                                        // 0 running, 1 finished OK, < 0 error (warp-uniform)
        if (n < 5) {                                                       // :414-418 (sic)
            uint32_t w = 0;
            if (n == 4) w = xpw_rd(in, 0, 4);
            st = (n == 0 || (n == 4 && w != 0xFFFFFFFFu)) ? 1 : ST_DATA;
        }
        while (st == 0) {
            if (ip + 4 > n) { st = ST_DATA; break; }                       // :461
            const uint32_t f = xpw_rd(in, ip, 4);
            ip += 4;
            const bool flagged = ((f >> (31 - lane)) & 1u) != 0;
            const uint32_t bsz = flagged ? 2u : 1u;
            uint32_t o = ip + (uint32_t)warp_incl_scan((int)bsz) - bsz;   // offset of my item, exact up to the first unresolved extension
            uint32_t mlen = 0, xbytes = 0;                                 // resolved extended length / its extension bytes (mine)
            unsigned resolved = 0;
            int xerr = 0;                                                  // error found while reading my extension
            uint32_t sym = 0;
            for (;;) {
                sym = (flagged && (uint64_t)o + 2 <= n) ? xpw_rd(in, o, 2) : 0u;
                const unsigned cand = __ballot_sync(FULL, flagged && (uint64_t)o + 2 <= n && (sym & 7u) == 7u) & ~resolved;
                if (!cand) break;
                const int e = __ffs((int)cand) - 1;                        // lowest unresolved extended match: its offset is exact
                uint32_t extra = 0, hv = half_val;
                bool hh = has_half;
                if (lane == e) {
                    uint32_t p = o + 2, len;
                    do {
                        if (hh) { len = hv >> 4; hh = false; }
                        else if (p >= n) { xerr = ST_DATA; len = 0; break; }
                        else { hv = xpw_rd(in, p, 1); hh = true; len = hv & 0xFu; ++p; }
                        if (len == 0xF) {
                            if (p >= n) { xerr = ST_DATA; break; }
                            len = xpw_rd(in, p, 1); ++p;
                            if (len == 0xFF) {
                                if ((uint64_t)p + 2 > n) { xerr = ST_DATA; break; }
                                len = xpw_rd(in, p, 2); p += 2;
                                if (len == 0) {
                                    if ((uint64_t)p + 4 > n) { xerr = ST_DATA; break; }
                                    len = xpw_rd(in, p, 4); p += 4;
                                }
                                if (len < 0xF + 0x7) { xerr = ST_DATA; break; }
                                len -= 0xF + 0x7;
                            }
                            len += 0xF;
                        }
                        len += 0x7;
                    } while (false);
                    mlen = len + 3;
                    extra = p - (o + 2);
                    xbytes = extra;
                }
                extra = __shfl_sync(FULL, extra, e);
                half_val = __shfl_sync(FULL, hv, e);
                has_half = __shfl_sync(FULL, (int)hh, e) != 0;
                const int ee = __shfl_sync(FULL, xerr, e);
                if (lane > e) o += extra;
                resolved |= 1u << e;
                if (ee) break;                                             // nothing behind a broken item is looked at
            }
            // what every lane's item is: end of stream, broken, literal or match
            const bool at_end = o == n;                                    // the stream ends where this item would start (:433-438)
            const unsigned endm = __ballot_sync(FULL, at_end);
            const int nitems = endm ? (__ffs((int)endm) - 1) : 32;
            const bool mine = lane < nitems;
            uint32_t len = 0, off = 0;
            int err = 0;
            if (mine) {
                if (flagged) {
                    if ((uint64_t)o + 2 > n) err = ST_DATA;
                    else if (xerr) err = xerr;
                    else { off = (sym >> 3) + 1; len = ((sym & 7u) == 7u) ? mlen : (sym & 7u) + 3; }
                } else len = 1;
            }
            const uint32_t pos = opos + (uint32_t)warp_incl_scan((int)len) - len;
            if (mine && !err) {
                if (flagged) { if (off > pos) err = ST_DATA; else if ((uint64_t)pos + len > cap) err = ST_BUF; }   // :442-443
                else if (pos == cap) err = ST_BUF;                                                                  // :455
            }
            int end_err = 0;
            if (endm && lane == nitems) end_err = ((f & (0xFFFFFFFFu >> lane)) == (0xFFFFFFFFu >> lane)) ? 0 : ST_DATA;
            const unsigned errm = __ballot_sync(FULL, err != 0 || end_err != 0);
            if (errm) { const int l = __ffs((int)errm) - 1; st = __shfl_sync(FULL, err ? err : end_err, l); break; }
            // tokens (a match longer than LZQ_MAXLEN takes several)
            const uint32_t nt = mine ? (flagged ? (len + LZQ_MAXLEN - 1) / LZQ_MAXLEN : 1u) : 0u;
            const uint32_t tincl = (uint32_t)warp_incl_scan((int)nt);
            if (mine) {
                uint32_t* t = tok + ntok + tincl - nt;
                if (!flagged) *t = lzt_lit(xpw_rd(in, o, 1));
                else { uint32_t l = len; while (l) { const uint32_t c = l > LZQ_MAXLEN ? LZQ_MAXLEN : l; *t++ = lzt_match(off, c); l -= c; } }
            }
            ntok += __shfl_sync(FULL, tincl, 31);
            opos = __shfl_sync(FULL, pos + len, 31 < nitems ? 31 : (nitems ? nitems - 1 : 0));
            if (endm) { st = 1; break; }
            ip = __shfl_sync(FULL, o + bsz + xbytes, 31);                  // first byte behind the last item = next flag word
        }
        if (lane == 0) { out_len[u] = opos; status[u] = st == 1 ? ST_OK : st; ntok_out[u] = ntok; }
    }
}

// Indexed parse: streams that come with a restart index (XpressSeg, made by xpress_encode_kernel) are parsed by one lane per
// SEGMENT instead of one lane per stream.  Lane i handles segment i % stride of stream i / stride; tokens of a segment go to
// tok + out_off + seg.out_pos (a token stands for at least one output byte, so segments cannot collide).
__global__ void __launch_bounds__(32 * XPD_WARPS_PER_BLOCK)
xpress_parse_indexed_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ tok, uint32_t count,
                            const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                            const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                            const XpressSeg* __restrict__ index, uint32_t stride,
                            uint32_t* __restrict__ seg_ntok, int32_t* __restrict__ seg_status, uint32_t* __restrict__ seg_end)
{
    const int lane = lane_id();
    const int wib = (int)(threadIdx.x >> 5);
    const uint32_t warps_total = gridDim.x * XPD_WARPS_PER_BLOCK;
    const uint64_t lanes = (uint64_t)count * stride;
    for (uint64_t g = blockIdx.x * XPD_WARPS_PER_BLOCK + wib; g * 32 < lanes; g += warps_total) {
        const uint64_t i = g * 32 + lane;
        const uint32_t u = (uint32_t)(i / stride), sgm = (uint32_t)(i % stride);
        XpLane L;
        L.in = in; L.in_len = 0; L.cap = 0; L.ip = 0; L.flags = 0; L.flagged = false; L.have_flags = false;
        L.half_val = 0; L.has_half = false; L.status = 1; L.stop_ip = 0xFFFFFFFFu;
        LaneTok O;
        O.init(tok);
        bool mine = false;
        if (i < lanes) {
            const XpressSeg* ix = index + (uint64_t)u * stride;
            uint32_t nseg = ix[0].nseg;
            if (nseg > stride) nseg = 0;                                    // not an index of this layout: the apply stage reports it
            if (sgm < nseg) {
                mine = true;
                const XpressSeg e = ix[sgm];
                L.in = in + in_off[u]; L.in_len = in_len[u]; L.cap = out_cap[u];
                L.status = 0;
                O.init(tok + out_off[u] + e.out_pos);
                O.opos = e.out_pos;
                if (sgm + 1 < nseg) L.stop_ip = ix[sgm + 1].in_pos;
                if (sgm == 0 && (e.in_pos != 0 || e.out_pos != 0 || e.flag_used != 0 || e.half_pos != 0xFFFFFFFFu)) {
                    L.status = ST_DATA;                                     // segment 0 is the start of the stream, nothing else
                    O.init(tok + out_off[u]); O.opos = 0;
                } else if (sgm == 0) {
                    if (L.in_len < 5) {                                     // :414-418 (sic)
                        uint32_t w = 0;
                        if (L.in_len == 4) w = (uint32_t)L.in[0] | ((uint32_t)L.in[1] << 8) | ((uint32_t)L.in[2] << 16) | ((uint32_t)L.in[3] << 24);
                        L.status = (L.in_len == 0 || (L.in_len == 4 && w != 0xFFFFFFFFu)) ? 1 : ST_DATA;
                    }
                } else if (e.in_pos > L.in_len || e.out_pos > L.cap || (e.flag_used && (e.flag_used > 31 || (uint64_t)e.flag_pos + 4 > L.in_len)) ||
                           (e.half_pos != 0xFFFFFFFFu && e.half_pos >= L.in_len)) {
                    L.status = ST_DATA;                                     // an index that does not belong to this stream
                } else {
                    L.ip = e.in_pos;
                    if (e.flag_used) {
                        const uint8_t* f = L.in + e.flag_pos;
                        const uint32_t fw = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
                        L.flagged = ((fw >> (31 - e.flag_used)) & 1u) != 0;
                        L.flags = (e.flag_used == 31 ? 0u : (fw << (e.flag_used + 1))) | (1u << e.flag_used);
                        L.have_flags = true;
                    }
                    if (e.half_pos != 0xFFFFFFFFu) { L.half_val = L.in[e.half_pos]; L.has_half = true; }
                }
            }
        }
        L.win.init(L.in, L.in + L.in_len);
        while (__any_sync(FULL, L.status == 0)) {
            if (L.status == 0) xp_step(L, O);
        }
        if (mine) { seg_ntok[i] = O.ntok; seg_status[i] = L.status == 1 ? ST_OK : L.status; seg_end[i] = O.opos; }
    }
}

// Second stage for indexed streams: one warp per STREAM executes the token runs of its segments in order (matches may reach
// back into earlier segments), after checking that every segment parsed cleanly and ended where the next one starts.
__global__ void __launch_bounds__(32 * LZA_WARPS_PER_BLOCK)
xpress_apply_indexed_kernel(const uint32_t* __restrict__ tok_base, const uint32_t* __restrict__ seg_ntok, const int32_t* __restrict__ seg_status,
                            const uint32_t* __restrict__ seg_end, const XpressSeg* __restrict__ index, uint32_t stride,
                            const uint64_t* __restrict__ out_off, uint32_t count, uint8_t* __restrict__ out,
                            uint32_t* __restrict__ out_len, int32_t* __restrict__ status)
{
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * LZA_WARPS_PER_BLOCK;
    for (uint32_t u = blockIdx.x * LZA_WARPS_PER_BLOCK + (threadIdx.x >> 5); u < count; u += warps_total) {
        const XpressSeg* ix = index + (uint64_t)u * stride;
        const uint64_t i0 = (uint64_t)u * stride;
        uint32_t nseg = ix[0].nseg;
        if (nseg > stride) nseg = 0;
        int bad = 0;                                                    // first failing segment's status
        for (uint32_t s0 = 0; s0 < nseg && !bad; s0 += 32) {
            const uint32_t sg = s0 + lane;
            int st = 0;
            if (sg < nseg) {
                st = seg_status[i0 + sg];
                if (st == 0 && sg + 1 < nseg && seg_end[i0 + sg] != ix[sg + 1].out_pos) st = ST_DATA;
            }
            const unsigned m = __ballot_sync(FULL, st < 0);
            if (m) bad = __shfl_sync(FULL, st, __ffs((int)m) - 1);
        }
        if (nseg == 0) bad = ST_DATA;
        if (bad) { if (lane == 0) { out_len[u] = 0; status[u] = bad; } continue; }
        uint8_t* o = out + out_off[u];
        for (uint32_t sg = 0; sg < nseg; ++sg)
            lz_apply_run(tok_base + out_off[u] + ix[sg].out_pos, seg_ntok[i0 + sg], o, ix[sg].out_pos);
        if (lane == 0) { out_len[u] = seg_end[i0 + nseg - 1]; status[u] = ST_OK; }
    }
}

// The same for few, large streams: one BLOCK per stream, the rounds of a segment pipelined over its warps (lz_apply.cuh).
__global__ void __launch_bounds__(32 * LZB_WARPS)
xpress_apply_indexed_block_kernel(const uint32_t* __restrict__ tok_base, const uint32_t* __restrict__ seg_ntok, const int32_t* __restrict__ seg_status,
                                  const uint32_t* __restrict__ seg_end, const XpressSeg* __restrict__ index, uint32_t stride,
                                  const uint64_t* __restrict__ out_off, uint32_t count, uint8_t* __restrict__ out,
                                  uint32_t* __restrict__ out_len, int32_t* __restrict__ status)
{
    __shared__ LzBlockSmem sm;
    __shared__ int s_bad;
    const int tid = (int)threadIdx.x;
    for (uint32_t u = blockIdx.x; u < count; u += gridDim.x) {
        const XpressSeg* ix = index + (uint64_t)u * stride;
        const uint64_t i0 = (uint64_t)u * stride;
        uint32_t nseg = ix[0].nseg;
        if (nseg > stride) nseg = 0;
        if (tid == 0) s_bad = 0x7fffffff;                               // first failing segment
        __syncthreads();
        for (uint32_t sg = tid; sg < nseg; sg += blockDim.x) {
            int st = seg_status[i0 + sg];
            if (st == 0 && sg + 1 < nseg && seg_end[i0 + sg] != ix[sg + 1].out_pos) st = ST_DATA;
            if (st < 0) atomicMin(&s_bad, (int)sg);
        }
        __syncthreads();
        int bad = 0;
        if (nseg == 0) bad = ST_DATA;
        else if (s_bad != 0x7fffffff) { bad = seg_status[i0 + (uint32_t)s_bad]; if (bad == 0) bad = ST_DATA; }
        if (bad) { if (tid == 0) { out_len[u] = 0; status[u] = bad; } __syncthreads(); continue; }
        uint8_t* o = out + out_off[u];
        for (uint32_t sg = 0; sg < nseg; ++sg)
            lz_apply_block_run(sm, tok_base + out_off[u] + ix[sg].out_pos, seg_ntok[i0 + sg], o, ix[sg].out_pos);
        if (tid == 0) { out_len[u] = seg_end[i0 + nseg - 1]; status[u] = ST_OK; }
        __syncthreads();
    }
}

} // namespace msc
