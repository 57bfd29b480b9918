// stream_logic.h -- the reference's LZNT1 stream state machine, independent of where the chunk codecs run.
//
// Replaces: lznt1_deflate_init / lznt1_deflate / lznt1_deflate_end ..... src/lznt1_compress.cpp:136-231 (+ chunk writer :95-135)
//           lznt1_inflate_init / lznt1_inflate / lznt1_inflate_end ..... src/lznt1_decompress.cpp:122-273
//           the copy streams of MSCOMP_NONE ............................... src/mscomp.cpp:26-61
//           stream bookkeeping (drain buffered output, top up the buffered partial chunk) ... include/mscomp/internal.h:472-520
// What a call consumes, produces and returns is the reference's, call by call.  The chunk codecs come in through `Codec`:
//   Codec::encode(in, n, comp, offs)    in[0..n) as LZNT1 chunks back to back (the last may be short), offs[c] = start of chunk c,
//                                       offs[nchunks] = total
//   Codec::decode(in, n, offs, out, sizes, status)   the chunks whose headers sit at in[offs[i]] -> out[4096 * i ..]
// The product instantiates it with the GPU batch codec (mscomp_stream.cu); the CPU test-suite instantiates the SAME logic with
// the oracle's chunk functions (tests/emu/stream_host.cpp) and drives it against the compiled reference.
#pragma once
#include "../../include/mscomp.h"
#include "../../include/lznt1.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace msc_stream {


constexpr size_t CHUNK = 4096;

struct CState {                       // lznt1_deflate
    bool finished;
    uint8_t in[CHUNK];  size_t in_needed, in_avail;       // a partial chunk waiting for more input
    uint8_t out[CHUNK + 2]; size_t out_pos, out_avail;    // the tail of a chunk that did not fit the output
};
struct DState {                       // lznt1_inflate
    bool end_of_stream;
    uint8_t in[CHUNK + 2]; size_t in_needed, in_avail;
    uint8_t out[CHUNK]; size_t out_pos, out_avail;
};

inline void set_error(mscomp_stream* s, const char* msg) { if (s) { strncpy(s->error, msg, sizeof(s->error) - 1); s->error[sizeof(s->error) - 1] = 0; } }
inline bool bad_stream(const mscomp_stream* s, bool compressing, MSCompFormat f)
{
    return !s || s->format != f || (bool)s->compressing != compressing || (!s->in && s->in_avail) || (!s->out && s->out_avail);
}
inline void init_stream(mscomp_stream* s, bool compressing, MSCompFormat f)
{
    s->format = f; s->compressing = compressing;
    s->in = nullptr; s->out = nullptr; s->in_avail = s->out_avail = 0; s->in_total = s->out_total = 0;
    s->error[0] = 0; s->warning[0] = 0; s->state = nullptr;
}
inline void adv_in(mscomp_stream* s, size_t n) { s->in += n; s->in_total += n; s->in_avail -= n; }
inline void adv_out(mscomp_stream* s, size_t n) { s->out += n; s->out_total += n; s->out_avail -= n; }

// Buffered output first (internal.h:472-492).  Returns true when the call has to stop (output full again).
template <class S> bool drain(S* st, mscomp_stream* s)
{
    if (!st->out_avail) return false;
    const size_t n = st->out_avail < s->out_avail ? st->out_avail : s->out_avail;
    memcpy(s->out, st->out + st->out_pos, n);
    s->out += n; s->out_total += n;
    if (st->out_avail == n) { s->out_avail -= n; st->out_avail = 0; return false; }
    s->out_avail = 0; st->out_pos += n; st->out_avail -= n;
    return true;
}

// ---- deflate --------------------------------------------------------------------------------------------------------
// One encoded chunk (header + payload = `bytes`, made from in_len input bytes) goes to the stream the way the reference's
// chunk writer sends it (src/lznt1_compress.cpp:95-135): straight to the output when a worst-case chunk would fit,
// through the state's buffer otherwise.
inline void emit_chunk(mscomp_stream* s, CState* st, const uint8_t* chunk, size_t bytes, size_t in_len)
{
    if (s->out_avail < in_len + 2) {
        memcpy(st->out, chunk, bytes);
        const size_t copy = bytes < s->out_avail ? bytes : s->out_avail;
        memcpy(s->out, st->out, copy);
        adv_out(s, copy);
        st->out_pos = copy;
        st->out_avail = bytes - copy;
    } else {
        memcpy(s->out, chunk, bytes);
        adv_out(s, bytes);
    }
}
template <class Codec>
MSCompStatus write_one(mscomp_stream* s, CState* st, const uint8_t* in, size_t in_len)
{
    std::vector<uint8_t> comp; std::vector<uint64_t> offs;
    const MSCompStatus r = Codec::encode(in, in_len, comp, offs);
    if (r != MSCOMP_OK) { set_error(s, "lznt1_deflate: the chunk encoder could not run"); return r; }
    emit_chunk(s, st, comp.data(), comp.size(), in_len);
    return MSCOMP_OK;
}

// ---- inflate --------------------------------------------------------------------------------------------------------
// A decoded compressed chunk (src/lznt1_decompress.cpp:153-186) / a stored chunk (:188-206) goes to the stream.
inline void emit_decoded(mscomp_stream* s, DState* st, const uint8_t* data, size_t out_size)
{
    if (s->out_avail < CHUNK) {
        memcpy(st->out, data, out_size);
        const size_t copy = out_size < s->out_avail ? out_size : s->out_avail;
        memcpy(s->out, st->out, copy);
        st->out_pos = copy;
        st->out_avail = out_size - copy;
        adv_out(s, copy);
    } else {
        memcpy(s->out, data, out_size);
        adv_out(s, out_size);
    }
}
inline void emit_stored(mscomp_stream* s, DState* st, const uint8_t* payload, size_t out_size)
{
    if (s->out_avail < out_size) {
        const size_t first = s->out_avail;
        memcpy(s->out, payload, first);
        memcpy(st->out, payload + first, out_size - first);
        st->out_pos = 0;
        st->out_avail = out_size - first;
        adv_out(s, first);
    } else {
        memcpy(s->out, payload, out_size);
        adv_out(s, out_size);
    }
}

// Header of the chunk at in[0..avail): what the reference's chunk reader decides before it touches the payload
// (src/lznt1_decompress.cpp:127-151).
enum HeaderKind { H_END, H_END_BAD, H_PARTIAL, H_BADSIG, H_COMPRESSED, H_STORED };
inline HeaderKind classify(const uint8_t* in, size_t avail, size_t total_left, size_t* in_size)
{
    const unsigned h = (unsigned)in[0] | ((unsigned)in[1] << 8);
    if (h == 0) return total_left == 2 ? H_END : H_END_BAD;
    *in_size = (h & 0x0FFF) + 3;
    if (*in_size > avail) return H_PARTIAL;
    if ((h & 0x7000) != 0x3000) return H_BADSIG;
    return (h & 0x8000) ? H_COMPRESSED : H_STORED;
}

// The reference's chunk reader for ONE chunk held in `in` (used for the chunk assembled in the state's buffer).
template <class Codec>
MSCompStatus read_one(mscomp_stream* s, DState* st, const uint8_t* in, size_t* in_len)
{
    size_t in_size = 0;
    switch (classify(in, *in_len, s->in_avail + st->in_avail, &in_size)) {
    case H_END: *in_len = 2; st->end_of_stream = true; return MSCOMP_OK;
    case H_END_BAD: set_error(s, "lznt1_inflate: data after the end-of-stream terminator"); return MSCOMP_DATA_ERROR;
    case H_PARTIAL:
        memmove(st->in, in, *in_len);
        st->in_needed = in_size - *in_len;
        st->in_avail = *in_len;
        return MSCOMP_OK;
    case H_BADSIG: set_error(s, "lznt1_inflate: chunk header without the 011 signature bits"); return MSCOMP_DATA_ERROR;
    case H_STORED: case H_COMPRESSED: break;
    }
    const HeaderKind kind = classify(in, *in_len, s->in_avail + st->in_avail, &in_size);
    *in_len = in_size;
    // (stored chunks go through the chunk decoder too -- on the GPU it copies them -- so that no chunk payload is handled on the CPU)
    std::vector<uint64_t> offs(1, 0); std::vector<uint8_t> out; std::vector<uint32_t> sizes; std::vector<int32_t> status;
    const MSCompStatus r = Codec::decode(in, in_size, offs, out, sizes, status);
    if (r != MSCOMP_OK) { set_error(s, "lznt1_inflate: the chunk decoder could not run"); return r; }
    if (status[0] < 0) { set_error(s, "lznt1_inflate: malformed chunk payload"); return MSCOMP_DATA_ERROR; }
    if (kind == H_STORED) emit_stored(s, st, out.data(), sizes[0]);
    else emit_decoded(s, st, out.data(), sizes[0]);
    return MSCOMP_OK;
}

inline bool possible_end(const mscomp_stream* s, const DState* st)
{
    return (!s->in_avail || (s->in_avail == 1 && s->in[0] == 0)) && (!st->in_avail || (st->in_avail == 1 && st->in[0] == 0)) && !st->out_avail;
}

// ---- MSCOMP_NONE -----------------------------------------------------------------------------------------------------
inline MSCompStatus copy_init(mscomp_stream* s) { if (!s) return MSCOMP_ARG_ERROR; init_stream(s, true, MSCOMP_NONE); return MSCOMP_OK; }
inline MSCompStatus copy_move(mscomp_stream* s)
{
    const size_t n = s->in_avail < s->out_avail ? s->in_avail : s->out_avail;
    if (n) memcpy(s->out, s->in, n);
    adv_out(s, n); adv_in(s, n);
    return MSCOMP_OK;
}
inline MSCompStatus copy_deflate(mscomp_stream* s, MSCompFlush flush)
{
    if (bad_stream(s, true, MSCOMP_NONE)) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    copy_move(s);
    return (flush == MSCOMP_FINISH && !s->in_avail) ? MSCOMP_STREAM_END : MSCOMP_OK;
}
inline MSCompStatus copy_inflate(mscomp_stream* s)
{   // (the reference initialises copy streams as "compressing" in both directions, src/mscomp.cpp:34,:48)
    if (bad_stream(s, true, MSCOMP_NONE)) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    return copy_move(s);
}
inline MSCompStatus copy_end(mscomp_stream* s)
{
    if (bad_stream(s, true, MSCOMP_NONE)) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    return MSCOMP_OK;
}


inline MSCompStatus lznt1_deflate_init_impl(mscomp_stream* s)
{
    if (!s) return MSCOMP_ARG_ERROR;
    init_stream(s, true, MSCOMP_LZNT1);
    CState* st = (CState*)malloc(sizeof(CState));
    if (!st) { set_error(s, "lznt1_deflate_init: out of memory for the stream state"); return MSCOMP_MEM_ERROR; }
    st->finished = false; st->in_needed = st->in_avail = 0; st->out_pos = st->out_avail = 0;
    s->state = (mscomp_internal_state*)st;
    return MSCOMP_OK;
}

inline MSCompStatus lznt1_deflate_end_impl(mscomp_stream* s)
{
    if (bad_stream(s, true, MSCOMP_LZNT1) || !s->state) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    CState* st = (CState*)s->state;
    MSCompStatus r = MSCOMP_OK;
    if (!st->finished || s->in_avail || st->in_avail || st->out_avail) { set_error(s, "lznt1_deflate_end: the stream was not finished (input or output still pending)"); r = MSCOMP_DATA_ERROR; }
    free(st);
    s->state = nullptr;
    return r;
}

inline MSCompStatus lznt1_inflate_init_impl(mscomp_stream* s)
{
    if (!s) return MSCOMP_ARG_ERROR;
    init_stream(s, false, MSCOMP_LZNT1);
    DState* st = (DState*)malloc(sizeof(DState));
    if (!st) { set_error(s, "lznt1_inflate_init: out of memory for the stream state"); return MSCOMP_MEM_ERROR; }
    st->end_of_stream = false; st->in_needed = st->in_avail = 0; st->out_pos = st->out_avail = 0;
    s->state = (mscomp_internal_state*)st;
    return MSCOMP_OK;
}

inline MSCompStatus lznt1_inflate_end_impl(mscomp_stream* s)
{
    if (bad_stream(s, false, MSCOMP_LZNT1) || !s->state) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    DState* st = (DState*)s->state;
    MSCompStatus r = MSCOMP_OK;
    if (!possible_end(s, st)) { set_error(s, "lznt1_inflate_end: the stream is in the middle of a chunk"); r = MSCOMP_DATA_ERROR; }
    free(st);
    s->state = nullptr;
    return r;
}

template <class Codec>
MSCompStatus lznt1_deflate_impl(mscomp_stream* s, MSCompFlush flush)
{
    CState* st = s ? (CState*)s->state : nullptr;
    if (bad_stream(s, true, MSCOMP_LZNT1) || !st || st->finished) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    MSCompStatus r;
    bool at_end = false;

    if (drain(st, s)) return MSCOMP_OK;
    if (st->in_avail) {                                        // top up the buffered partial chunk (:175-195)
        const size_t copy = st->in_needed < s->in_avail ? st->in_needed : s->in_avail;
        if (copy) { memcpy(st->in + st->in_avail, s->in, copy); st->in_avail += copy; st->in_needed -= copy; adv_in(s, copy); }
        if (st->in_needed) {
            if (flush != MSCOMP_NO_FLUSH) {                    // a flush cuts the chunk short
                if ((r = write_one<Codec>(s, st, st->in, st->in_avail)) != MSCOMP_OK) return r;
                st->in_avail = 0; st->in_needed = 0;
                if (flush == MSCOMP_FINISH && !st->out_avail) at_end = true;
            }
            if (!at_end) return MSCOMP_OK;
        } else {
            if ((r = write_one<Codec>(s, st, st->in, CHUNK)) != MSCOMP_OK) return r;
            st->in_avail = 0;
        }
    }
    if (!at_end) {
        // Full chunks while there is room in the output (:198-202).  All the chunks this call can get rid of are encoded as
        // one GPU batch; how many fit is only known afterwards (a chunk takes at most 4098 bytes), so a batch is sized for
        // the worst case plus one and the loop runs again if the output still has room.
        while (s->out_avail && s->in_avail >= CHUNK) {
            size_t want = s->out_avail / (CHUNK + 2) + 1;
            if (want > s->in_avail / CHUNK) want = s->in_avail / CHUNK;
            std::vector<uint8_t> comp; std::vector<uint64_t> offs;
            if ((r = Codec::encode(s->in, want * CHUNK, comp, offs)) != MSCOMP_OK) { set_error(s, "lznt1_deflate: the chunk encoder could not run"); return r; }
            for (size_t c = 0; c < want && s->out_avail; ++c) {
                emit_chunk(s, st, comp.data() + offs[c], (size_t)(offs[c + 1] - offs[c]), CHUNK);
                adv_in(s, CHUNK);
            }
        }
        if (s->out_avail && s->in_avail) {                     // 0 < in_avail < 4096 (:205-219)
            if (flush != MSCOMP_NO_FLUSH) {
                if ((r = write_one<Codec>(s, st, s->in, s->in_avail)) != MSCOMP_OK) return r;
            } else {
                memcpy(st->in, s->in, s->in_avail);
                st->in_avail = s->in_avail;
                st->in_needed = CHUNK - st->in_avail;
            }
            adv_in(s, s->in_avail);
        }
        at_end = flush == MSCOMP_FINISH && !s->in_avail && !st->in_avail && !st->out_avail;
    }
    if (at_end) {
        st->finished = true;
        if (s->out_avail >= 2) { s->out[0] = s->out[1] = 0; }   // end-of-buffer terminal, not counted (:225-227)
        return MSCOMP_STREAM_END;
    }
    return MSCOMP_OK;
}

template <class Codec>
MSCompStatus lznt1_inflate_impl(mscomp_stream* s)
{
    if (bad_stream(s, false, MSCOMP_LZNT1) || !s->state) { set_error(s, "stream argument is null, of another format / direction, or has a null buffer with a non-zero size"); return MSCOMP_ARG_ERROR; }
    DState* st = (DState*)s->state;
    MSCompStatus r;

    if (drain(st, s)) return MSCOMP_OK;
    if (st->end_of_stream) {
        if (s->in_avail || st->in_avail) { set_error(s, "lznt1_inflate: data after the end-of-stream terminator"); return MSCOMP_DATA_ERROR; }
        return MSCOMP_STREAM_END;
    }
    if (st->in_avail) {                                        // the chunk being assembled in the state's buffer (:236-245)
        for (;;) {
            const size_t copy = st->in_needed < s->in_avail ? st->in_needed : s->in_avail;
            if (copy) { memcpy(st->in + st->in_avail, s->in, copy); st->in_avail += copy; st->in_needed -= copy; adv_in(s, copy); }
            if (st->in_needed) return MSCOMP_OK;
            size_t in_len = st->in_avail;
            r = read_one<Codec>(s, st, st->in, &in_len);
            if (in_len != st->in_avail) { set_error(s, "lznt1_inflate: buffered chunk does not match its header"); return MSCOMP_ARG_ERROR; }
            if (r != MSCOMP_OK) return r;
            if (st->end_of_stream) return MSCOMP_STREAM_END;
            if (st->in_needed) continue;                       // there was only enough for the header so far
            break;
        }
        st->in_avail = 0;
    }

    // Whole chunks while there is room in the output (:248-255).  The headers are walked first; the compressed chunks of the
    // run that can be processed are decoded as one GPU batch, then the run is replayed in order.
    while (s->out_avail && s->in_avail >= 2) {
        struct Item { size_t off, in_size; HeaderKind kind; int slot; };
        std::vector<Item> run;
        std::vector<uint64_t> offs;
        const size_t room = s->out_avail / CHUNK + 1;           // chunks that would fit if every one were full, plus one
        size_t p = 0;
        while (run.size() < room && s->in_avail - p >= 2) {
            Item it{p, 0, H_END, -1};
            it.kind = classify(s->in + p, s->in_avail - p, s->in_avail - p + st->in_avail, &it.in_size);
            if (it.kind == H_COMPRESSED || it.kind == H_STORED) { it.slot = (int)offs.size(); offs.push_back(p); }   // (stored chunks are copied by the GPU decoder)
            run.push_back(it);
            if (it.kind != H_COMPRESSED && it.kind != H_STORED) break;      // the replay stops here
            p += it.in_size;
        }
        std::vector<uint8_t> out; std::vector<uint32_t> sizes; std::vector<int32_t> status;
        if (!offs.empty() && (r = Codec::decode(s->in, p, offs, out, sizes, status)) != MSCOMP_OK) { set_error(s, "lznt1_inflate: the chunk decoder could not run"); return r; }
        for (size_t i = 0; i < run.size() && s->out_avail && s->in_avail >= 2; ++i) {
            const Item& it = run[i];
            const uint8_t* in = s->in;                          // == start of the run + it.off (everything before was consumed)
            (void)in;
            switch (it.kind) {
            case H_END: adv_in(s, 2); st->end_of_stream = true; return MSCOMP_STREAM_END;
            case H_END_BAD: set_error(s, "lznt1_inflate: data after the end-of-stream terminator"); return MSCOMP_DATA_ERROR;
            case H_BADSIG: set_error(s, "lznt1_inflate: chunk header without the 011 signature bits"); return MSCOMP_DATA_ERROR;
            case H_PARTIAL:                                     // the rest of the input is the beginning of a chunk
                memcpy(st->in, in, s->in_avail);
                st->in_needed = it.in_size - s->in_avail;
                st->in_avail = s->in_avail;
                adv_in(s, s->in_avail);
                break;
            case H_STORED:
                if (status[it.slot] < 0) { set_error(s, "lznt1_inflate: malformed chunk payload"); return MSCOMP_DATA_ERROR; }
                emit_stored(s, st, out.data() + (size_t)it.slot * CHUNK, sizes[it.slot]);
                adv_in(s, it.in_size);
                break;
            case H_COMPRESSED:
                if (status[it.slot] < 0) { set_error(s, "lznt1_inflate: malformed chunk payload"); return MSCOMP_DATA_ERROR; }
                emit_decoded(s, st, out.data() + (size_t)it.slot * CHUNK, sizes[it.slot]);
                adv_in(s, it.in_size);
                break;
            }
        }
    }
    if (s->out_avail && s->in_avail) {                          // one byte left: half a header (:257-264)
        st->in[0] = s->in[0];
        st->in_needed = 1;
        st->in_avail = 1;
        s->in += 1; s->in_total += 1; s->in_avail = 0;
    }
    return possible_end(s, st) ? MSCOMP_POSSIBLE_STREAM_END : MSCOMP_OK;
}


// ---- plain Xpress: xpress_inflate (src/xpress_decompress.cpp:45-61, :247-403) ----------------------------------------------------
// The decoder state and the step that advances it live where the codec runs (kernels/xpress_stream.cuh); `Backend` says where
// that is: the GPU (mscomp_stream.cu) or, for the CPU test-suite, host memory (tests/emu/stream_host.cpp).
//   Backend::create() -> state handle or nullptr      Backend::destroy(handle)
//   Backend::step(handle, in, in_len, out, out_cap, &in_used, &out_made, &status) -> MSCompStatus of the transport
//   Backend::tail(handle, &flags, &flagged, &kept)     what xpress_inflate_end looks at
template <class Backend> MSCompStatus xpress_inflate_init_impl(mscomp_stream* s)
{
    if (!s) return MSCOMP_ARG_ERROR;
    init_stream(s, false, MSCOMP_XPRESS);
    void* h = Backend::create();
    if (!h) { set_error(s, "xpress_inflate_init: no memory for the decoder state"); return MSCOMP_MEM_ERROR; }
    s->state = (mscomp_internal_state*)h;
    return MSCOMP_OK;
}
template <class Backend> MSCompStatus xpress_inflate_impl(mscomp_stream* s)
{
    if (bad_stream(s, false, MSCOMP_XPRESS) || !s->state) { set_error(s, "xpress_inflate: not an open Xpress decompression stream"); return MSCOMP_ARG_ERROR; }
    for (;;) {
        // one step handles up to 1 GiB on either side; a larger call is a sequence of steps with the same outcome
        const size_t kMax = (size_t)1 << 30;
        const uint32_t n_in = (uint32_t)(s->in_avail < kMax ? s->in_avail : kMax), n_out = (uint32_t)(s->out_avail < kMax ? s->out_avail : kMax);
        uint32_t used = 0, made = 0; int st = 0;
        const MSCompStatus t = Backend::step(s->state, s->in, n_in, s->out, n_out, &used, &made, &st);
        if (t != MSCOMP_OK) { set_error(s, "xpress_inflate: the device did not complete the step"); return t; }
        adv_in(s, used); adv_out(s, made);
        if (st < 0) { set_error(s, "xpress_inflate: the data is not a valid Xpress stream"); return (MSCompStatus)st; }
        const bool clipped = (n_in == kMax && s->in_avail) || (n_out == kMax && s->out_avail);
        if (!clipped || (used == 0 && made == 0)) return (MSCompStatus)st;
    }
}
template <class Backend> MSCompStatus xpress_inflate_end_impl(mscomp_stream* s)
{
    if (bad_stream(s, false, MSCOMP_XPRESS) || !s->state) { set_error(s, "xpress_inflate_end: not an open Xpress decompression stream"); return MSCOMP_ARG_ERROR; }
    uint32_t flags = 0, flagged = 0, kept = 0;
    Backend::tail(s->state, &flags, &flagged, &kept);
    MSCompStatus st = MSCOMP_OK;
    const uint32_t inv = ~flags;
    const bool ones_on_top = ((inv + 1u) & inv) == 0;
    if (s->in_avail || kept || !flagged || !ones_on_top) { set_error(s, "xpress_inflate_end: the stream has not reached a possible end"); st = MSCOMP_DATA_ERROR; }   // :394
    Backend::destroy(s->state);
    s->state = nullptr;
    return st;
}
// xpress_deflate*: upstream never finished them -- the release build answers MSCOMP_MEM_ERROR from init, MSCOMP_ARG_ERROR from
// deflate, and deflate_end only checks its argument (src/xpress_compress.cpp:50-74, :75-221, :223-236).  Same answers here.
inline MSCompStatus xpress_deflate_init_impl(mscomp_stream*) { return MSCOMP_MEM_ERROR; }
inline MSCompStatus xpress_deflate_impl(mscomp_stream*, MSCompFlush) { return MSCOMP_ARG_ERROR; }
inline MSCompStatus xpress_deflate_end_impl(mscomp_stream* s)
{
    if (bad_stream(s, true, MSCOMP_XPRESS) || !s->state) { set_error(s, "xpress_deflate_end: not an open Xpress compression stream"); return MSCOMP_ARG_ERROR; }
    return MSCOMP_OK;
}

} // namespace msc_stream
