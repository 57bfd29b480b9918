// xpress_stream.cuh -- resumable plain-Xpress decoder behind xpress_inflate (the streaming API).
//
// Replaces the state machine of the reference's xpress_inflate (src/xpress_decompress.cpp:247-403, symbol reader :62-113,
// history include/mscomp/CircularBuffer.h).  A call decodes whatever the caller handed in: it stops when the input is used up
// -- possibly in the middle of a flag word or of a symbol, whose bytes are then kept for the next call -- or when the output
// is full -- possibly in the middle of a match, whose remainder is then owed to the next call.  What a call consumes,
// produces and returns equals the reference's, call by call (tests: the differential driver of tests/stream_util.py).
//
// The stream is byte-serial (flag words, shared length nibbles, an 8 KiB sliding history), and a call may be a single byte
// long, so there is no parallel unit here: ONE thread walks the items; the state (including the 8 KiB history) lives in device
// memory between calls.  This is API completeness, not a throughput path -- bulk data belongs to
// xpress_decompress / the batch API.
#pragma once
#ifdef MSC_HOST_ONLY            // the CPU test-suite compiles the same logic with a plain host compiler (tests/emu/stream_host.cpp)
#include <stdint.h>
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
namespace msc { enum : int { ST_OK = 0, ST_ARG = -2, ST_DATA = -3, ST_MEM = -4, ST_BUF = -5 }; }
#else
#include "common.cuh"
#endif

namespace msc {

constexpr uint32_t XPS_HIST = 0x2000;                 // history = window of the format (offsets are 1..8192)

struct XpStreamState {
    uint32_t flags, flagged;                          // flag word in force (shifted as items are consumed) / kind of the next item
    uint32_t has_half, half_byte;                     // a byte whose high length nibble is still unused
    uint32_t stash_n;                                 // input bytes kept from the previous call (a partial flag word or symbol)
    uint8_t stash[16];
    uint32_t copy_len, copy_off;                      // the rest of a match that did not fit the output
    uint32_t hist_size, hist_pos;                     // bytes in the history (<= XPS_HIST) / next write slot
    uint32_t upstream_quirk;                          // test hook: reproduce the reference's bookkeeping slip (see xps_step, step A)
    uint8_t hist[XPS_HIST];
};
struct XpStreamResult { uint32_t in_used, out_made; int32_t status; uint32_t pad_; };

constexpr int XPS_POSSIBLE_END = 2;                   // MSCOMP_POSSIBLE_STREAM_END

__device__ __forceinline__ void xps_init(XpStreamState* s)
{
    s->flags = 0; s->flagged = 1; s->has_half = 0; s->half_byte = 0; s->stash_n = 0; s->copy_len = 0; s->copy_off = 0;
    s->hist_size = 0; s->hist_pos = 0; s->upstream_quirk = 0;
}
// all set bits above all clear bits (also: none set, all set) -- the shape of a flag word whose remaining items are "end of stream"
__device__ __forceinline__ bool xps_ones_on_top(uint32_t x) { x = ~x; return ((x + 1) & x) == 0; }

__device__ __forceinline__ void xps_push(XpStreamState* s, uint8_t b)
{
    s->hist[s->hist_pos] = b;
    s->hist_pos = (s->hist_pos + 1) & (XPS_HIST - 1);
    if (s->hist_size < XPS_HIST) ++s->hist_size;
}
// n bytes from distance off of the history to out (and into the history), byte-serial semantics
__device__ __forceinline__ void xps_copy(XpStreamState* s, uint32_t off, uint32_t n, uint8_t* out)
{
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t b = s->hist[(s->hist_pos - off) & (XPS_HIST - 1)];
        out[i] = b;
        xps_push(s, b);
    }
}

// Symbol at p[0..avail) (:62-113).  Returns 0 and sets used/off/len, or 1 = not enough bytes (adv = bytes taken so far without
// the nibble byte, half_updated = the half-byte state changed), or a negative status.  half / has_half are updated in place.
__device__ __forceinline__ int xps_symbol(const uint8_t* p, uint32_t avail, uint32_t& used, uint32_t& off, uint32_t& len,
                                          uint32_t& has_half, uint32_t& half_byte, bool& half_from_here, uint32_t& adv, bool& half_updated)
{
    uint32_t i = 0;
    adv = 0; half_updated = false; half_from_here = false;
    if (avail < 2) return 1;
    const uint32_t sym = (uint32_t)p[0] | ((uint32_t)p[1] << 8);
    i = 2;
    off = (sym >> 3) + 1; len = sym & 7u;
    if (len == 7u) {
        if (has_half) { len = half_byte >> 4; has_half = 0; }
        else if (i == avail) { adv = 2; return 1; }
        else { half_byte = p[i++]; has_half = 1; half_from_here = true; len = half_byte & 15u; }
        if (len == 15u) {
            if (i == avail) { adv = 2; half_updated = true; return 1; }
            len = p[i++];
            if (len == 0xFFu) {
                if (i + 2 > avail) { adv = 3; half_updated = true; return 1; }
                len = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8); i += 2;
                if (len == 0) {
                    if (i + 4 > avail) { adv = 5; half_updated = true; return 1; }
                    len = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24); i += 4;
                }
                if (len < 22u) return ST_DATA;                            // :104
                len -= 22u;
            }
            len += 15u;
        }
        len += 7u;
    }
    len += 3u;
    used = i;
    return 0;
}

// One call of xpress_inflate on device-resident buffers.
__device__ __forceinline__ void xps_step(XpStreamState* s, const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_cap, XpStreamResult* res)
{
    uint32_t i = 0, o = 0;
    uint32_t i_keep = 0, o_keep = 0;                                       // what an error inside the item loop reports: the reference commits
                                                                           // its positions only when a call ends regularly
    bool looping = false;
    auto done = [&](int st) {
        const bool err = st < 0 && looping;
        res->in_used = err ? i_keep : i; res->out_made = err ? o_keep : o; res->status = st;
    };
    // A. the rest of a match owed from the previous call (:255-270).  When it still does not fit, the reference subtracts the
    // output room AFTER having zeroed it (:261-263: ADVANCE_OUT_TO_END, then copy_len -= stream->out_avail), so a match owed
    // over three or more calls comes out too long there -- corrupt output.  The owed length is kept right here; the quirk flag
    // reproduces the reference's arithmetic so that the call-by-call differential tests can also run such schedules.
    if (s->copy_len) {
        if (s->copy_len > out_cap) {
            xps_copy(s, s->copy_off, out_cap, out); o = out_cap;
            if (!s->upstream_quirk) s->copy_len -= out_cap;
            return done(ST_OK);
        }
        xps_copy(s, s->copy_off, s->copy_len, out);
        o = s->copy_len;
        s->copy_len = 0;
        s->flagged = s->flags & 0x80000000u;
        s->flags <<= 1;
    }
    uint32_t flags = s->flags, flagged = s->flagged, off = 0, len = 0;
    uint32_t has_half = s->has_half, half_byte = s->half_byte;
    int start = 0;                                                         // 1 = inside a flag word, 2 = a symbol is read, 3 = a literal is due
    bool half_in_stash = false;
    auto save = [&]() { s->flags = flags; s->flagged = flagged; s->has_half = has_half; s->half_byte = half_byte; };
    // B. where the previous call stopped (:279-323)
    if (!flags) {
        if (s->stash_n) {                                                  // a flag word cut in two
            if (in_len + s->stash_n < 4) {
                for (uint32_t k = 0; k < in_len; ++k) s->stash[s->stash_n + k] = in[k];
                s->stash_n += in_len; i = in_len;
                return done(ST_OK);
            }
            const uint32_t take = 4 - s->stash_n;
            for (uint32_t k = 0; k < take; ++k) s->stash[s->stash_n + k] = in[k];
            i = take;
            const uint32_t w = (uint32_t)s->stash[0] | ((uint32_t)s->stash[1] << 8) | ((uint32_t)s->stash[2] << 16) | ((uint32_t)s->stash[3] << 24);
            flagged = w & 0x80000000u;
            flags = (w << 1) | 1u;
            s->stash_n = 0;
            start = 1;
        }
    } else if (flagged) {                                                  // a symbol, possibly cut: up to 10 bytes are put together
        uint32_t to_copy = 10 - s->stash_n;
        if (to_copy > in_len) to_copy = in_len;
        for (uint32_t k = 0; k < to_copy; ++k) s->stash[s->stash_n + k] = in[k];
        uint32_t used = 0, adv = 0; bool from_here = false, hupd = false;
        uint32_t hh = has_half, hb = half_byte;
        const int r = xps_symbol(s->stash, s->stash_n + to_copy, used, off, len, hh, hb, from_here, adv, hupd);
        if (r < 0) { return done(r); }
        if (r == 1) {                                                      // still not whole: everything offered is kept (:230-234)
            i = to_copy; s->stash_n += to_copy;
            const bool maybe_end = adv == 0 && i == in_len && s->stash_n == 0 && xps_ones_on_top(s->flags);
            return done(maybe_end ? XPS_POSSIBLE_END : ST_OK);
        }
        has_half = hh; half_byte = hb;
        half_in_stash = from_here;                                         // (the reference keeps a POINTER to the nibble byte: here, into its stash)
        i = used - s->stash_n;
        s->stash_n = 0;
        start = 2;
    } else {
        if (!in_len) return done(ST_OK);
        start = 3;
    }
    // C. the item loop (:337-397)
    i_keep = i; o_keep = o; looping = true;
    for (;;) {
        if (start == 0) {
            if (i + 4 > in_len) break;
            const uint32_t w = (uint32_t)in[i] | ((uint32_t)in[i + 1] << 8) | ((uint32_t)in[i + 2] << 16) | ((uint32_t)in[i + 3] << 24);
            flagged = w & 0x80000000u;
            flags = (w << 1) | 1u;
            i += 4;
        }
        do {
            if (start == 0 || start == 1) {
                start = 0;
                if (i == in_len) { save(); return done(flagged && xps_ones_on_top(flags) ? XPS_POSSIBLE_END : ST_OK); }
            }
            if (start == 2 || (start == 0 && flagged)) {
                if (start == 0) {
                    uint32_t used = 0, adv = 0; bool from_here = false, hupd = false;
                    uint32_t hh = has_half, hb = half_byte;
                    const int r = xps_symbol(in + i, in_len - i, used, off, len, hh, hb, from_here, adv, hupd);
                    if (r < 0) { return done(r); }
                    if (r == 1) {
                        // not whole: the symbol's bytes are kept for the next call together with a matching half-byte state (:236-253)
                        uint32_t keep_from = i;                            // (the reader has not moved i)
                        if (!hupd) { /* half-byte state as it was before the symbol */ }
                        else if (from_here) { has_half = 0; }              // the nibble byte was taken by this symbol: it goes back too
                        else { has_half = 1; half_byte = 0xF0u; }          // the high nibble of an earlier byte was used: it read 15
                        s->stash_n = in_len - keep_from;
                        for (uint32_t k = 0; k < s->stash_n; ++k) s->stash[k] = in[keep_from + k];
                        i = in_len;
                        save();
                        const bool maybe_end = adv == 0 && s->stash_n == 0 && xps_ones_on_top(flags);
                        return done(maybe_end ? XPS_POSSIBLE_END : ST_OK);
                    }
                    if (hh != has_half || hb != half_byte) half_in_stash = false;
                    has_half = hh; half_byte = hb;
                    i += used;
                }
                start = 0;
                if (s->hist_size < off) return done(ST_DATA);             // :369
                const uint32_t room = out_cap - o;
                if (len > room) {
                    xps_copy(s, off, room, out + o);
                    o = out_cap;
                    s->copy_len = len - room; s->copy_off = off;
                    save();
                    return done(ST_OK);
                }
                xps_copy(s, off, len, out + o);
                o += len;
            } else {
                start = 0;
                if (o == out_cap) { save(); return done(ST_OK); }
                const uint8_t b = in[i++];
                out[o++] = b;
                xps_push(s, b);
            }
            flagged = flags & 0x80000000u;
            flags <<= 1;
        } while (flags);
    }
    // fewer than four bytes are left: they wait for the rest of their flag word (:398-402)
    s->stash_n = in_len - i;
    for (uint32_t k = 0; k < s->stash_n; ++k) s->stash[k] = in[i + k];
    // (second slip of the reference, reproduced only under the test hook: it saves the half byte through its pointer AFTER this
    // copy, so a nibble byte that sat at stash[2] is read back as whatever the copy put there)
    if (s->upstream_quirk && half_in_stash && has_half && s->stash_n > 2) half_byte = s->stash[2];
    i = in_len;
    save();
    return done(ST_OK);
}

#ifndef MSC_HOST_ONLY
__global__ void xpress_stream_init_kernel(XpStreamState* s, uint32_t quirk) { if (threadIdx.x == 0 && blockIdx.x == 0) { xps_init(s); s->upstream_quirk = quirk; } }
__global__ void xpress_stream_step_kernel(XpStreamState* s, const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_cap, XpStreamResult* res)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) xps_step(s, in, in_len, out, out_cap, res);
}
#endif

} // namespace msc
