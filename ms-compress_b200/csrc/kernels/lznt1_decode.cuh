// lznt1_decode.cuh -- LZNT1 chunk decoder, one warp per 4 KiB chunk.
//
// Replaces the reference's lznt1_decompress_chunk (src/lznt1_decompress.cpp:37-121) and the per-chunk part of
// lznt1_decompress_chunk_read (:122-209: header parse, signature check, stored-chunk copy).
//
// Wire format: u16 header = (payload bytes - 1) | 0x3000 | 0x8000 if compressed; compressed
// payload = { flag byte (LSB first, 1 = token); up to 8 items: literal byte or u16 token }.  A token at chunk
// output position p splits as off = (tok >> shift) + 1, len = (tok & ((1 << shift) - 1)) + 3 with
// shift = 12 while p <= 16 and one less each time p passes a power of two (reference :59-63).
//
// Data movement: the payload is brought into shared memory by ONE bulk asynchronous copy (cp.async.bulk + mbarrier)
// and the decoded chunk leaves by one bulk store; nothing passes through registers.  Decoding is IN PLACE: the
// payload sits right-aligned at the top of a 4096+512-byte buffer and the output grows from byte 0.  A group of n
// input bytes yields at least n-1 output bytes, so in a chunk that ends within 4096 bytes the write position never
// reaches the unread input (margin: 4096/9 < 512); a chunk where it would is one that overflows 4096 bytes, i.e. an
// error in the reference as well (:67, :99) -- the round that would cross the line reports it.
//
// Parallel scheme, per round of up to 32 flag groups (lane i owns group i):
//   1. chase: start(g+1) = start(g) + 9 + popc(flags(g)) does not depend on token contents; lane i simply takes i
//      steps of that chain (predicated, no cross-lane traffic; bytes read past the payload are harmless slack).
//   2. the 8 items of the group go to registers together with the group's output length under the shift in force
//      at the round's first byte; one warp scan (length and token count packed in one word) gives every group its
//      output position and its first slot in the round's match list.
//   3. if a token of the round starts beyond the current power-of-two limit, lengths are re-evaluated with the
//      shift each token has at its own (tentative) position and re-scanned until nothing moves: positions are
//      exact from the left, every pass fixes everything up to and including the next boundary crossing.
//   4. literals are stored; tokens are appended, in position order, to a per-warp list {token, position, shift}.
//   5. the list is executed 32 entries at a time: entries whose source lies wholly below the first entry of the
//      batch (everything there is final) and that are short and not self-overlapping are copied at once, one per
//      lane; the few others (a source inside the batch, a run, a long match) follow in order, copied by the whole
//      warp.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int LZNT1_CHUNK = 4096;
constexpr int LZNT1_WARPS_PER_BLOCK = 4;
constexpr int LZD_BUF = LZNT1_CHUNK + 512;        // in-place buffer: output from 0, payload right-aligned below LZD_BUF
constexpr int LZD_SLACK = 544;                    // readable bytes behind it: the chase may run 31 * 17 bytes past the payload
constexpr int LZD_LIST = 256;                     // tokens per round: 32 groups x 8 items
constexpr int LZD_BULK_LEN = 16;                  // matches up to this long are copied one per lane

struct Lznt1WarpSmem {
    __align__(16) uint8_t buf[LZD_BUF + LZD_SLACK];
    uint32_t list[LZD_LIST];                      // token | position << 16 | (shift - 4) << 28
    __align__(8) uint64_t bar;                    // mbarrier of the payload load
    uint64_t pad_;
};
static_assert(sizeof(Lznt1WarpSmem) % 16 == 0, "per-warp shared memory must keep 16-byte alignment");

// shift / limit in force for a token that starts at chunk output position p
__device__ __forceinline__ void lznt1_shift_limit(int p, int& shift, int& limit)
{
    const int m = (p <= 16) ? 4 : (32 - __clz(p - 1));
    shift = 16 - m;
    limit = 1 << m;
}
// the same for positions that may be garbage (a damaged stream): the shift never leaves [4, 12]
__device__ __forceinline__ int lzd_token_shift(int p)
{
    int x = p - 1;
    x = x < 0 ? 0 : x;
    int m = 32 - __clz(x);                           // p <= 1 gives 0
    m = m < 4 ? 4 : m;
    m = m > 12 ? 12 : m;
    return 16 - m;
}

// Whole-warp copy of one match inside the buffer with the semantics of a byte-serial copy (all arguments uniform).
__device__ __forceinline__ void lzd_coop_copy(uint8_t* o, int pos, int off, int len)
{
    const int lane = lane_id();
    int src = pos - off;
    if (off < 32 && off < len) {
        // A run: the output repeats the `off` bytes before pos.  The first stretch is indexed modulo the period; after
        // that the copy distance can be any multiple of the period, so one of at least 32 is used and 32-byte batches
        // never read what they write.
        const int off2 = off * ((31 + off) / off);
        const int head = len < off2 ? len : off2;
        for (int i = lane; i < head; i += 32) o[pos + i] = o[src + i % off];
        __syncwarp();
        pos += head; len -= head;
        src = pos - off2;
    }
    for (int i = 0; i < len; i += 32) {
        if (i + lane < len) o[pos + i + lane] = o[src + i + lane];
        __syncwarp();
    }
}

// Group length with the shift every token has at its own position, the group starting at gs (items in registers).
// ONE: the round crosses at most one power of two (positions stay <= 2 * lim): the shift is S0 up to lim and S1 = S0 - 1 behind it.
template <bool ALL8, bool ONE>
__device__ __forceinline__ int lzd_group_len(const uint32_t (&it)[8], unsigned f, int nit, int gs, int S0, int S1, int lim)
{
    int pos = gs;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (ALL8 || j < nit) {
            const int Sj = ONE ? (pos > lim ? S1 : S0) : lzd_token_shift(pos);
            const int l = (int)(it[j] & ((1u << Sj) - 1)) + 3;
            pos += ((f >> j) & 1) ? l : 1;
        }
    }
    const int g = pos - gs;
    return g < LZNT1_CHUNK + 1 ? g : LZNT1_CHUNK + 1;
}

// Literals to the buffer, tokens to the match list; returns true if a token reaches back beyond the chunk start.
// MIXED: 0 = every token of the round has shift S0; 1 = one power of two is crossed (S0 up to lim, S1 behind it); 2 = any.
template <bool ALL8, int MIXED>
__device__ __forceinline__ bool lzd_emit(const uint32_t (&it)[8], unsigned f, int nit, int gs, int S0, int S1, int lim, int li, uint8_t* buf, uint32_t* list)
{
    bool bad = false;
    int pos = gs;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (ALL8 || j < nit) {
            const uint32_t t = it[j];
            const bool tk = (f >> j) & 1;
            const int Sj = MIXED == 2 ? lzd_token_shift(pos) : (MIXED == 1 ? (pos > lim ? S1 : S0) : S0);
            const int l = (int)(t & ((1u << Sj) - 1)) + 3, off = (int)(t >> Sj) + 1;
            bad |= tk && off > pos;                                          // :65 / :98
            if (tk) list[li] = t | ((uint32_t)pos << 16) | ((uint32_t)(Sj - 4) << 28);
            else buf[pos] = (uint8_t)t;
            li += tk ? 1 : 0;
            pos += tk ? l : 1;
        }
    }
    return bad;
}

// Decodes the compressed payload buf[in0 .. in0+n) into buf[0 ..); returns bytes produced or a negative status.
__device__ __forceinline__ int lznt1_decode_payload(Lznt1WarpSmem& sm, const int in0, const int n)
{
    uint8_t* const buf = sm.buf;
    const uint32_t buf_sa = smem_u32(sm.buf);
    const uint8_t* const pin = sm.buf + in0;
    const int lane = lane_id();
    int q = 0;          // payload offset of the round's first flag byte (uniform)
    int P = 0;          // output bytes produced so far (uniform)

    while (q < n) {
        // ---- 1. chase: lane i walks i groups ahead of q ------------------------------------------------------
        uint32_t sa = buf_sa + (uint32_t)(in0 + q);                          // (shared address; 4 instructions per step: test, load, popc, add)
#pragma unroll
        for (int step = 0; step < 31; ++step)
            if (step < lane) sa += 9u + (uint32_t)__popc(lds_u8(buf, buf_sa, sa));
        const int s = (int)(sa - buf_sa) - in0;
        const bool act = s < n;
        const unsigned f = act ? (unsigned)pin[s] : 0u;
        const int qn = __shfl_sync(FULL, s + 9 + __popc(f), 31);            // first flag byte of the next round (>= n: none)

        // ---- 2. items to registers, group length under the round's first shift -------------------------------
        const int S0 = lzd_token_shift(P);
        const int lim = 1 << (16 - S0);                                      // last position that still has shift S0
        const unsigned mask0 = (1u << S0) - 1;
        uint32_t it[8];
        int nit = 0, A = 0, lastA = -1;                                      // items present, length, start of my last token (relative)
        bool trunc = false;                                                  // a token cut by the end of the payload (:90)
        const bool all8 = __all_sync(FULL, s + 17 <= n);                     // every group of the round lies wholly inside the payload
        if (all8) {
            int o = s + 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t v = pin[o];
                if ((f >> j) & 1) { v |= (uint32_t)pin[o + 1] << 8; lastA = A; A += (int)(v & mask0) + 3; o += 2; }
                else { A += 1; o += 1; }
                it[j] = v;
            }
            nit = 8;
        } else {
            int o = s + 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t v = 0;
                if (act && o < n) {                                          // (the stream may end after any item, :86)
                    if ((f >> j) & 1) {
                        if (o + 2 > n) { trunc = true; o = n; }
                        else { v = (uint32_t)pin[o] | ((uint32_t)pin[o + 1] << 8); lastA = A; A += (int)(v & mask0) + 3; o += 2; nit = j + 1; }
                    } else { v = pin[o]; A += 1; o += 1; nit = j + 1; }
                }
                it[j] = v;
            }
        }
        const int ntk = __popc(f & ((1u << nit) - 1u));
        __syncwarp();                                                        // all input of the round is in registers: output may overwrite it

        // ---- 3. positions ---------------------------------------------------------------------------------------
        int glen = A < LZNT1_CHUNK + 1 ? A : LZNT1_CHUNK + 1;
        int incl = warp_incl_scan(glen | (ntk << 20));
        int gs = P + (incl & 0xFFFFF) - glen;
        int mixed = 0;
        const int S1 = S0 > 4 ? S0 - 1 : 4;
        if (__any_sync(FULL, lastA >= 0 && gs + lastA > lim)) {
            // Positions taken under S0 can only be too large.  If even they stay within the next power of two, every token
            // behind the crossing has shift S0 - 1 at its tentative and at its true position alike: one pass settles the round.
            const bool single = P + (__shfl_sync(FULL, incl, 31) & 0xFFFFF) <= 2 * lim;
            mixed = single ? 1 : 2;
            for (int iter = 0; iter < 34; ++iter) {
                const int g2 = single ? (all8 ? lzd_group_len<true, true>(it, f, nit, gs, S0, S1, lim) : lzd_group_len<false, true>(it, f, nit, gs, S0, S1, lim))
                                      : (all8 ? lzd_group_len<true, false>(it, f, nit, gs, S0, S1, lim) : lzd_group_len<false, false>(it, f, nit, gs, S0, S1, lim));
                const bool moved = g2 != glen;
                glen = g2;
                incl = warp_incl_scan(glen | (ntk << 20));
                gs = P + (incl & 0xFFFFF) - glen;
                if (single || !__any_sync(FULL, moved)) break;               // lengths agree with the positions they were taken at
            }
        }
        const int total = __shfl_sync(FULL, incl, 31);
        const int Pend = P + (total & 0xFFFFF);
        const int nm = total >> 20;
        const bool cut = __any_sync(FULL, trunc);
        if (cut || Pend > LZNT1_CHUNK || (qn < n && Pend > in0 + qn)) return ST_DATA;   // :90, :67/:99 (see the header on the last one)

        // ---- 4. literals out, tokens to the list -----------------------------------------------------------------
        const int li = (incl >> 20) - ntk;
        bool bad;
        if (all8) bad = mixed == 0 ? lzd_emit<true, 0>(it, f, nit, gs, S0, S1, lim, li, buf, sm.list)
                      : (mixed == 1 ? lzd_emit<true, 1>(it, f, nit, gs, S0, S1, lim, li, buf, sm.list) : lzd_emit<true, 2>(it, f, nit, gs, S0, S1, lim, li, buf, sm.list));
        else bad = mixed == 0 ? lzd_emit<false, 0>(it, f, nit, gs, S0, S1, lim, li, buf, sm.list)
                 : (mixed == 1 ? lzd_emit<false, 1>(it, f, nit, gs, S0, S1, lim, li, buf, sm.list) : lzd_emit<false, 2>(it, f, nit, gs, S0, S1, lim, li, buf, sm.list));
        if (__any_sync(FULL, bad)) return ST_DATA;
        __syncwarp();

        // ---- 5. matches --------------------------------------------------------------------------------------------
        for (int base = 0; base < nm; base += 32) {
            const bool have = base + lane < nm;
            const uint32_t e = have ? sm.list[base + lane] : 0u;
            const int Sj = (int)(e >> 28) + 4, pos = (int)(e >> 16) & 0xFFF, t = (int)(e & 0xFFFFu);
            const int len = (t & ((1 << Sj) - 1)) + 3, off = (t >> Sj) + 1, src = pos - off;
            const int W0 = __shfl_sync(FULL, pos, 0);                        // everything below the batch's first match is final
            const bool now = have && len <= LZD_BULK_LEN && off >= len && src + len <= W0;
            if (now) {
                buf[pos] = buf[src]; buf[pos + 1] = buf[src + 1]; buf[pos + 2] = buf[src + 2];
                for (int i = 3; i < len; ++i) buf[pos + i] = buf[src + i];
            }
            __syncwarp();
            unsigned rest = __ballot_sync(FULL, have && !now);
            while (rest) {
                const int o = __ffs((int)rest) - 1;
                rest &= rest - 1;
                lzd_coop_copy(buf, __shfl_sync(FULL, pos, o), __shfl_sync(FULL, off, o), __shfl_sync(FULL, len, o));
            }
        }
        P = Pend;
        q = qn;
    }
    __syncwarp();
    return P;
}

// One warp per chunk.  chunk_off[c] = byte offset of chunk c's header inside `in`.
// Chunk c is written to out + c*4096 (slot layout); out_size[c] = bytes produced, status[c] = 0 or < 0.
// `in` must be readable up to the next 16-byte boundary behind in_len (the bulk copy moves whole 16-byte granules).
__global__ void __launch_bounds__(LZNT1_WARPS_PER_BLOCK * 32, 8)
lznt1_decode_kernel(const uint8_t* __restrict__ in, uint64_t in_len, const uint64_t* __restrict__ chunk_off, uint32_t nchunks,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ out_size, int32_t* __restrict__ status)
{
    __shared__ Lznt1WarpSmem smem[LZNT1_WARPS_PER_BLOCK];
    const int lane = lane_id();
    const int wib = (int)(threadIdx.x >> 5);
    Lznt1WarpSmem& sm = smem[wib];
    const uint32_t warps_total = gridDim.x * LZNT1_WARPS_PER_BLOCK;
    if (lane == 0) mbar_init(&sm.bar, 1);
    __syncwarp();
    uint32_t phase = 0;

    for (uint32_t c = blockIdx.x * LZNT1_WARPS_PER_BLOCK + wib; c < nchunks; c += warps_total) {
        const uint64_t start = chunk_off[c];
        uint8_t* dst = out + (uint64_t)c * LZNT1_CHUNK;
        int st = ST_OK;
        uint32_t produced = 0;
        if (start + 2 > in_len) { st = ST_BUF; }
        else {
            const uint32_t hdr = (uint32_t)in[start] | ((uint32_t)in[start + 1] << 8);
            const uint32_t n = (hdr & 0x0FFF) + 1;                       // payload bytes
            if (hdr == 0) { st = ST_DATA; }                              // terminator is handled by the caller's index walk
            else if (start + 2 + n > in_len) { st = ST_BUF; }            // truncated chunk (reference: internal.h:627)
            else if ((hdr & 0x7000) != 0x3000) { st = ST_DATA; }         // src/lznt1_decompress.cpp:152
            else if (!(hdr & 0x8000)) {                                  // stored chunk :188-206
                warp_copy_bytes(dst, in + start + 2, n);
                produced = n;
            } else {
                // payload -> top of the buffer: the 16-byte granules covering [src, src+n), one bulk copy
                const uint8_t* src = in + start + 2;
                const uint32_t skew = (uint32_t)((uintptr_t)src & 15u);
                const uint32_t wbytes = (skew + n + 15u) & ~15u;
                if (lane == 0) {
                    bulk_store_wait_read();                              // the previous chunk has left the buffer
                    fence_async_smem();
                    bulk_load_g2s(sm.buf + LZD_BUF - wbytes, src - skew, wbytes, &sm.bar);
                }
                mbar_wait(&sm.bar, phase);
                phase ^= 1u;
                const int r = lznt1_decode_payload(sm, (int)(LZD_BUF - wbytes + skew), (int)n);
                if (r < 0) st = r;
                else {
                    produced = (uint32_t)r;
                    fence_async_smem();                                  // the bytes written above, as seen by the bulk store
                    __syncwarp();
                    uint32_t nb = 0;
                    if (((uintptr_t)dst & 15u) == 0) {
                        nb = produced & ~15u;
                        if (lane == 0 && nb) bulk_store_s2g(dst, sm.buf, nb);
                    }
                    for (uint32_t b = nb + lane; b < produced; b += 32) dst[b] = sm.buf[b];
                }
                __syncwarp();
            }
        }
        if (lane == 0) { out_size[c] = produced; status[c] = st; }
    }
    if (lane == 0) bulk_store_wait_read();
}

} // namespace msc
