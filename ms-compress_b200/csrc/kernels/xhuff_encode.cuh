// xhuff_encode.cuh -- Xpress-Huffman encoder, one 256-thread block per buffer (64 KiB chunk after chunk).
//
// Replaces the reference's xpress_huff_compress (src/xpress_huff_compress.cpp:247-331): LZ77 + symbol histogram
// (xh_compress_lz77 :52-153), code construction (HuffmanEncoder::CreateCodes, include/mscomp/HuffmanEncoder.h:58-127),
// the expansion guard with its literal-only fallback (:274-280, :310-316, xh_compress_no_matching :155-180) and the
// bit packer (xh_compress_encode :195-245 over OutputBitstream, include/mscomp/Bitstream.h:111-148).
//
// Differences from the reference that keep the stream valid for its decoder: the match
// window is reset at every 64 KiB chunk, unused symbols get code length 0, code lengths come from a plain
// Huffman tree clamped to 15 bits by a Kraft-sum repair instead of the reference's rescale-and-retry.
//
// Stages per chunk:  tiles of 2048 positions -> tokens (lz_front.cuh) + 512-bin histogram (shared atomics)
//   -> code lengths / canonical codes -> exact size (expansion guard) -> 256-byte table
//   -> pass A over tokens: running bit / raw-byte prefix sums give every code its bit address, every raw length
//      byte and every 16-bit word its final byte address (placement model of SURVEY.md A.3, which mirrors
//      OutputBitstream: two words are always reserved ahead, raw bytes go to the write cursor)
//   -> pass B: the assembled 16-bit words are scattered to their addresses.
//
// Restart index (optional, side-band; xhuff_index.cuh): pass A knows, for every token, how many bits and raw bytes precede it,
// and that is all the reference's reader state at that symbol depends on -- words fetched F = 2 at the start, else
// 1 + ceil(bits / 16) (Bitstream.h:59-70: a refill whenever fewer than 16 bits are left), valid bits = 16 F - bits, byte cursor
// = 2 F + raw bytes.  So the encoder writes the entry of every 64th symbol itself, identical to what the serial walk
// (xhuff_index_kernel) finds, and the decoder can start with one lane per segment without any walk.
#pragma once
#include "common.cuh"
#include "lz_front.cuh"
#include "lz_exact.cuh"
#include "xhuff_index.cuh"

namespace msc {

constexpr int XHE_THREADS = LZF_THREADS;
constexpr uint32_t XHE_CHUNK = 65536;
constexpr int XHE_SYMS = 512;
constexpr uint32_t XHE_TOK_CAP = XHE_CHUNK + 64;                 // tokens per chunk (+EOS)
constexpr uint32_t XHE_WORD_CAP = 36864;                         // 16-bit words per chunk: <= 1 + ceil(bits/16)
// per-block global scratch layout (bytes)
constexpr size_t XHE_SCR_TOK = 0;
constexpr size_t XHE_SCR_BITS = XHE_SCR_TOK + (size_t)XHE_TOK_CAP * 4;
constexpr size_t XHE_SCR_WPOS = XHE_SCR_BITS + (size_t)XHE_WORD_CAP * 2;
constexpr size_t XHE_SCR_SEGPOS = XHE_SCR_WPOS + (size_t)XHE_WORD_CAP * 4;      // input position of every 64th token (restart index)
constexpr size_t XHE_SCRATCH_BYTES = XHE_SCR_SEGPOS + ((size_t)XHE_TOK_CAP / XHI_SEG_SYMS + 2) * 4;

struct XhEncSmem {
    LzFrontSmem lz;
    uint32_t hist[XHE_SYMS];
    // Scratch of the code builder and the emitter.  It lives in the hash table of the LZ77 front end, which is dead once the
    // chunk's tokens exist (and is cleared again before the next chunk): 9.5 KB less shared memory = one more block per SM.
    __device__ __forceinline__ uint32_t* key() { return reinterpret_cast<uint32_t*>(lz.H); }                   // [XHE_SYMS] (count << 9 | symbol), sorted ascending
    __device__ __forceinline__ uint32_t* w() { return reinterpret_cast<uint32_t*>(lz.H) + XHE_SYMS; }           // [2 * XHE_SYMS] node weights
    __device__ __forceinline__ uint16_t* parent() { return reinterpret_cast<uint16_t*>(lz.H) + 6 * XHE_SYMS; }  // [2 * XHE_SYMS]
    __device__ __forceinline__ uint16_t* code() { return reinterpret_cast<uint16_t*>(lz.H) + 8 * XHE_SYMS; }    // [XHE_SYMS]
    __device__ __forceinline__ uint8_t* len() { return reinterpret_cast<uint8_t*>(lz.H) + 18 * XHE_SYMS; }      // [XHE_SYMS]
    uint32_t blc[16];                  // codes per length
    uint32_t nused, raw_total, bits_total, scan_tmp[8];
    unsigned long long bit_carry;      // running (bits << 32 | raw bytes) between token tiles
    unsigned long long wsum[XHE_THREADS / 32];   // per-warp partial sums of the block scans
    uint32_t big_pos, red;             // reference-parse variant: the tile's visited match of LZX_CAP or more bytes (LZF_TILE = none)
};

// Token word of this encoder: a match is (len - 1) << 16 | offset (len 3..65536, offset 1..65535, so the high half is at least 2);
// a literal is its byte value; the end-of-stream symbol is XHE_TOK_EOS (high half 1: neither of the two).
constexpr uint32_t XHE_TOK_EOS = 0x00010000u;
__device__ __forceinline__ bool xhe_is_match(uint32_t tok) { return (tok >> 16) >= 2u; }
__device__ __forceinline__ uint32_t xhe_sym_of(uint32_t tok)
{
    if (!xhe_is_match(tok)) return (tok == XHE_TOK_EOS) ? 256u : (tok & 0xFFu);
    const uint32_t off = tok & 0xFFFFu, lm3 = (tok >> 16) + 1 - 3;
    const uint32_t ob = 31u - (uint32_t)__clz((int)off);
    return 256u + ((ob << 4) | min(lm3, 15u));
}
__device__ __forceinline__ uint32_t xhe_raw_of(uint32_t tok)
{   // raw length bytes after the code: 0, 1 or 3 (a match ends with its 64 KiB chunk, the u32 form never occurs)
    if (!xhe_is_match(tok)) return 0;
    const uint32_t lm3 = (tok >> 16) + 1 - 3;
    return lm3 < 15 ? 0u : (lm3 < 15 + 255 ? 1u : 3u);
}

// Code lengths (<= 15) and canonical codes from sm.hist.  All threads.
__device__ __forceinline__ void xhe_build_codes(XhEncSmem& sm)
{
    const int tid = (int)threadIdx.x;
    for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) { sm.key()[i] = (sm.hist[i] << 9) | (uint32_t)i; sm.len()[i] = 0; sm.code()[i] = 0; }
    if (tid < 16) sm.blc[tid] = 0;
    __syncthreads();
    // bitonic sort of 512 keys, ascending
    for (int k = 2; k <= XHE_SYMS; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = sm.key()[i], b = sm.key()[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { sm.key()[i] = b; sm.key()[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        int first = 0;
        while (first < XHE_SYMS && (sm.key()[first] >> 9) == 0) ++first;
        const int m = XHE_SYMS - first;                            // used symbols, ascending by (count, symbol)
        sm.nused = (uint32_t)m;
        if (m == 1) { sm.parent()[0] = 0; }
        else if (m > 1) {                                          // two-queue Huffman: leaves 0..m-1, internal m..2m-2
            for (int i = 0; i < m; ++i) sm.w()[i] = sm.key()[first + i] >> 9;
            int i = 0, j = m, k = m;
            while (k < 2 * m - 1) {
                int a, b;
                if (i < m && (j >= k || sm.w()[i] <= sm.w()[j])) a = i++; else a = j++;
                if (i < m && (j >= k || sm.w()[i] <= sm.w()[j])) b = i++; else b = j++;
                sm.w()[k] = sm.w()[a] + sm.w()[b];
                sm.parent()[a] = (uint16_t)k; sm.parent()[b] = (uint16_t)k;
                ++k;
            }
        }
    }
    __syncthreads();
    const int m = (int)sm.nused, first = XHE_SYMS - m;
    // depth of every leaf, clamped to 15; codes per length
    for (int i = tid; i < m; i += XHE_THREADS) {
        int d = 0;
        if (m == 1) d = 1;
        else for (int x = i; x != 2 * m - 2; x = sm.parent()[x]) ++d;
        if (d > 15) d = 15;
        atomicAdd(&sm.blc[d], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        // Kraft repair: while over-subscribed, lengthen one code of the longest length below 15
        uint32_t K = 0;
        for (int l = 1; l <= 15; ++l) K += sm.blc[l] << (15 - l);
        while (K > 32768u) {
            int l = 14;
            while (sm.blc[l] == 0) --l;
            sm.blc[l]--; sm.blc[l + 1]++;
            K -= 1u << (14 - l);
        }
        // lengths by rank: the least frequent symbols take the longest codes
        int idx = 0;
        for (int l = 15; l >= 1; --l)
            for (uint32_t c = 0; c < sm.blc[l]; ++c) sm.len()[sm.key()[first + idx++] & 511] = (uint8_t)l;
        // canonical codes, ascending (length, symbol)  (HuffmanEncoder.h:115-123)
        uint32_t next[16], code = 0;
        for (int l = 1; l <= 15; ++l) { next[l] = code; code = (code + sm.blc[l]) << 1; }
        for (int s = 0; s < XHE_SYMS; ++s) { const int l = sm.len()[s]; if (l) sm.code()[s] = (uint16_t)next[l]++; }
    }
    __syncthreads();
}

// The reference's code lengths, tie for tie: HuffmanEncoder::CreateCodes (include/mscomp/HuffmanEncoder.h:58-127) is bzip2's
// heap-driven Huffman construction over ALL 512 symbols (an unused symbol weighs 1, so every symbol gets a code), weights carry the
// subtree depth in their low byte, and while some code is longer than 15 bits all weights are halved (+1) and the tree is rebuilt.
// Which of several equally heavy nodes is merged first -- and with it which of two equally frequent symbols gets the longer code --
// is decided by the order the binary heap happens to hold them in, so the heap is replayed operation for operation: one thread,
// heap entries that carry their weight (one shared-memory load per comparison).  The depths are then read off by all threads.
// Scratch: the parse arrays of the LZ77 front end (dead while codes are built).
__device__ __forceinline__ void xhe_build_codes_reference(XhEncSmem& sm)
{
    const int tid = (int)threadIdx.x;
    unsigned long long* heap = reinterpret_cast<unsigned long long*>(sm.lz.nxt);                  // [514]  weight << 32 | node
    uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(sm.lz.nxt) + 4128);     // [1024] node weights (count << 8 | depth)
    uint16_t* parent = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(sm.lz.nxt) + 8224); // [1024]
    static_assert(8224 + 2048 <= 3 * LZF_TILE * sizeof(uint16_t), "scratch of the code builder fits the parse arrays");
    for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) { const uint32_t c = sm.hist[i]; w[i + 1] = (c ? c : 1u) << 8; }
    if (tid == 0) w[0] = 0;
    __syncthreads();
    for (;;) {
        for (int i = tid; i < 2 * XHE_SYMS; i += XHE_THREADS) parent[i] = 0;
        __syncthreads();
        if (tid == 0) {
            uint32_t hl = 0;
            heap[0] = 0;                                                      // node 0, weight 0: the sentinel above the root
            // The replay is one long chain of dependent shared-memory loads, so both primitives fetch everything a step can need in
            // one go: a push all the ancestors of the new leaf (their indices are known up front), a pop both children AND the four
            // grandchildren (the sift only rewrites ancestors, so what was fetched stays true) -- two levels per round trip.
            auto push = [&](uint32_t x, uint32_t wx) {
                uint32_t j = ++hl;
                unsigned long long up[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) up[k] = heap[j >> (k + 1)];             // j <= 513: the 10th ancestor is the sentinel at 0
#pragma unroll
                for (int k = 0; k < 10; ++k) { if (!(wx < (uint32_t)(up[k] >> 32))) break; heap[j] = up[k]; j >>= 1; }
                heap[j] = ((unsigned long long)wx << 32) | x;
            };
            auto wof = [](unsigned long long e2) -> uint32_t { return (uint32_t)(e2 >> 32); };
            auto pop = [&]() {
                const unsigned long long t = heap[hl--];
                const uint32_t wt = (uint32_t)(t >> 32);
                uint32_t i = 1;
                for (;;) {
                    const uint32_t j = i << 1;
                    if (j > hl) break;
                    const unsigned long long c0 = heap[j], c1 = heap[min(j + 1u, 513u)];
                    const unsigned long long g0 = heap[min(2u * j, 513u)], g1 = heap[min(2u * j + 1u, 513u)], g2 = heap[min(2u * j + 2u, 513u)], g3 = heap[min(2u * j + 3u, 513u)];
                    const bool right = j < hl && wof(c1) < wof(c0);
                    const unsigned long long c = right ? c1 : c0;
                    if (wt < wof(c)) break;
                    heap[i] = c; i = right ? j + 1u : j;
                    const uint32_t jj = i << 1;
                    if (jj > hl) break;
                    const unsigned long long ga = right ? g2 : g0, gb = right ? g3 : g1;
                    const bool right2 = jj < hl && wof(gb) < wof(ga);
                    const unsigned long long g = right2 ? gb : ga;
                    if (wt < wof(g)) break;
                    heap[i] = g; i = right2 ? jj + 1u : jj;
                }
                heap[i] = t;
            };
            for (uint32_t i = 1; i <= (uint32_t)XHE_SYMS; ++i) push(i, w[i]);
            uint32_t nn = XHE_SYMS;
            while (hl > 1) {
                const unsigned long long e1 = heap[1]; pop();
                const unsigned long long e2 = heap[1]; pop();
                const uint32_t n1 = (uint32_t)e1, n2 = (uint32_t)e2, w1 = (uint32_t)(e1 >> 32), w2 = (uint32_t)(e2 >> 32);
                ++nn;
                parent[n1] = (uint16_t)nn; parent[n2] = (uint16_t)nn;
                const uint32_t wn = ((w1 & 0xffffff00u) + (w2 & 0xffffff00u)) | (1u + max(w1 & 0xffu, w2 & 0xffu));
                w[nn] = wn;
                push(nn, wn);
            }
        }
        __syncthreads();
        int too_long = 0;
        for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) {
            uint32_t d = 0;
            for (uint32_t k = (uint32_t)i + 1; parent[k]; k = parent[k]) ++d;
            sm.len()[i] = (uint8_t)min(d, 255u);
            if (d > 15) too_long = 1;
        }
        if (tid == 0) sm.scan_tmp[0] = 0;
        __syncthreads();
        if (too_long) sm.scan_tmp[0] = 1;
        __syncthreads();
        const bool again = sm.scan_tmp[0] != 0;
        __syncthreads();
        if (!again) break;
        for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) w[i + 1] = (1u + (w[i + 1] >> 9)) << 8;
        __syncthreads();
    }
    // canonical codes, ascending (length, symbol) (:104-116); every symbol has a length here
    if (tid < 16) sm.blc[tid] = 0;
    __syncthreads();
    for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) atomicAdd(&sm.blc[sm.len()[i]], 1u);
    __syncthreads();
    if (tid == 0) {
        uint32_t next[16], code = 0;
        for (int l = 1; l <= 15; ++l) { next[l] = code; code = (code + sm.blc[l]) << 1; }
        for (int s2 = 0; s2 < XHE_SYMS; ++s2) { const int l = sm.len()[s2]; sm.code()[s2] = (uint16_t)next[l]++; }
        sm.nused = XHE_SYMS;
    }
    __syncthreads();
}

__device__ __forceinline__ void xhe_or16(uint32_t* bits32, uint32_t j, uint32_t v)
{   // OR a 16-bit value into word j of the little-endian 16-bit word array
    atomicOr(&bits32[j >> 1], (v & 0xFFFFu) << (16 * (j & 1)));
}
// write the n-bit value v at bit address b of the MSB-first bit string
__device__ __forceinline__ void xhe_put_bits(uint32_t* bits32, uint32_t b, uint32_t v, uint32_t n)
{
    if (!n) return;
    const uint32_t j = b >> 4, avail = 16 - (b & 15);
    if (n <= avail) xhe_or16(bits32, j, v << (avail - n));
    else { xhe_or16(bits32, j, v >> (n - avail)); xhe_or16(bits32, j + 1, v << (16 - (n - avail))); }
}
__device__ __forceinline__ uint32_t xhe_reserved(uint32_t B) { return B ? 1 + ((B + 15) >> 4) : 2; }   // words reserved at B bits

// The reference's LZ77 for this encoder (EXACT): the answers of lz_find_kernel (window 65535, clipped at the 64 KiB chunk), the assumed
// parse per tile of lz_parse0_kernel, indexed by find coordinates = position in the STREAM (eos_mode 1: the units are its chunks, unit u
// starts at u * 65536, seg_first unused) or in the unit (eos_mode 0, segment base seg_first[u]).
struct XhPre {
    const uint32_t* seg_first;
    const uint32_t* lenoff;
    const uint32_t* visited0;
    const uint32_t* exit0;
    const uint16_t* dist;                      // the links, for questions nobody has asked yet
    const uint32_t* stream_len;                // eos_mode 1: length of the whole stream (device word)
};

// One block per buffer.  eos_mode 0: every buffer is its own stream (ends with the end-of-stream symbol);
// eos_mode 1: the buffers are consecutive chunks of ONE stream -- only the last buffer of the batch carries it.
// EXACT = false: lz_front.cuh's bulk-synchronous matcher, window reset at every chunk (valid streams, not the reference's tokens);
// EXACT = true: the reference's tokens (xh_compress_lz77, src/xpress_huff_compress.cpp:52-153) from lz_exact.cuh.
template <bool EXACT>
__device__ __forceinline__ void xhuff_encode_body(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                    const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                    const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                    uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint8_t* __restrict__ scratch, int eos_mode,
                    XhSeg* __restrict__ index, const uint64_t* __restrict__ index_off, uint32_t* __restrict__ nseg, const XhPre pre)
{
    MSC_DYN_SMEM(smem_raw);
    XhEncSmem& sm = *reinterpret_cast<XhEncSmem*>(smem_raw);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* scr = scratch + (size_t)blockIdx.x * XHE_SCRATCH_BYTES;
    uint32_t* tok = reinterpret_cast<uint32_t*>(scr + XHE_SCR_TOK);
    uint32_t* bits32 = reinterpret_cast<uint32_t*>(scr + XHE_SCR_BITS);
    uint16_t* bits16 = reinterpret_cast<uint16_t*>(scr + XHE_SCR_BITS);
    uint32_t* wpos = reinterpret_cast<uint32_t*>(scr + XHE_SCR_WPOS);
    uint32_t* segpos = reinterpret_cast<uint32_t*>(scr + XHE_SCR_SEGPOS);

    for (uint32_t u = blockIdx.x; u < count; u += gridDim.x) {
        const uint8_t* ubase = in + in_off[u];
        const uint32_t ulen = in_len[u], cap = out_cap[u];
        uint8_t* obase = out + out_off[u];
        uint32_t opos = 0;
        int st = ST_OK;
        const bool stream_end_here = (eos_mode == 0) || (u + 1 == count);
        XhSeg* ix = index ? index + index_off[u] : nullptr;              // this stream's restart index (eos_mode 0 only)
        const uint64_t ix_room = index ? index_off[u + 1] - index_off[u] : 0;
        uint32_t nent = 0;                                               // entries of the chunks before this one
        // find coordinates of this unit (EXACT): source, length and first position of the thing lz_find_kernel saw
        const uint8_t* fsrc = (eos_mode == 1) ? in + in_off[0] : ubase;
        const uint32_t fn = (EXACT && eos_mode == 1) ? *pre.stream_len : ulen;
        const uint32_t fpos0 = (eos_mode == 1) ? u * XHE_CHUNK : 0u;
        const uint64_t fbase = (uint64_t)((EXACT && eos_mode != 1) ? pre.seg_first[u] : u) * XHE_CHUNK;      // index of the unit's position 0 in lenoff[]

        for (uint32_t c0 = 0; c0 < ulen && st == ST_OK; c0 += XHE_CHUNK) {
            const uint8_t* src = ubase + c0;
            const uint32_t n = min(ulen - c0, XHE_CHUNK);
            const bool last = (c0 + n == ulen) && stream_end_here;

            // ---- LZ77 -> tokens + histogram ----
            if (!EXACT) for (int i = tid; i < (1 << LZF_HBITS); i += XHE_THREADS) sm.lz.H[i] = 0;
            for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) sm.hist[i] = 0;
            if (tid == 0) { sm.lz.next_p = 0; sm.lz.ntok = 0; }
            __syncthreads();
            bool literal_only = false;
            for (int attempt = 0; attempt < 2; ++attempt) {
                if (!literal_only) {
                    for (uint32_t t0 = 0; t0 < n; t0 += LZF_TILE) {
                        uint32_t cnt;
                        if (EXACT) {
                            const uint32_t* lo = pre.lenoff + fbase + c0 + t0;
                            const uint64_t gt = (fbase + c0 + t0) / LZF_TILE;
                            for (int r = tid; r < LZF_TILE; r += XHE_THREADS) sm.lz.lenoff[r] = (t0 + (uint32_t)r < n) ? lo[r] : 0u;
                            if (tid < LZF_TILE / 32) sm.lz.V[tid] = pre.visited0[gt * (LZF_TILE / 32) + tid];
                            if (tid == 0) sm.big_pos = LZF_TILE;
                            __syncthreads();
                            lzx_merge_parse(sm.lz, n, t0, pre.exit0[gt] - (fpos0 + c0), 0u, LzxUnit(fsrc, fn, pre.dist + fbase - fpos0, 65535u, XHE_CHUNK),
                                            fpos0 + c0 + t0);
                            // a visited match of LZX_CAP bytes or more leaves the tile, so there is at most one: measure it
                            for (int r = tid; r < LZF_TILE; r += XHE_THREADS)
                                if (((sm.lz.V[r >> 5] >> (r & 31)) & 1) && (sm.lz.lenoff[r] >> 16) == LZX_CAP - 1u) sm.big_pos = (uint32_t)r;
                            __syncthreads();
                            const uint32_t bp = sm.big_pos;
                            if (bp < (uint32_t)LZF_TILE) {
                                const uint32_t P = fpos0 + c0 + t0 + bp, off = sm.lz.lenoff[bp] & 0xFFFFu;
                                const uint32_t len = lzx_extend(fsrc, P, off, min(fn - P - 1u, XHE_CHUNK - (P % XHE_CHUNK)), &sm.red);
                                if (tid == 0) { sm.lz.lenoff[bp] = ((len - 1u) << 16) | off; sm.lz.next_p = t0 + bp + len; }
                                __syncthreads();
                            }
                            cnt = sm.lz.prefV[LZF_TILE / 32 - 1] + (uint32_t)__popc(sm.lz.V[LZF_TILE / 32 - 1]);
                        } else cnt = lzf_tile(sm.lz, src, n, t0, 65535u, 4096u);
                        const uint32_t tb = sm.lz.ntok;
                        for (int r = tid; r < LZF_TILE; r += XHE_THREADS) {
                            const uint32_t v = sm.lz.V[r >> 5];
                            if ((v >> (r & 31)) & 1) {
                                const uint32_t k = tb + sm.lz.prefV[r >> 5] + (uint32_t)__popc(v & ((1u << (r & 31)) - 1));
                                const uint32_t lo = sm.lz.lenoff[r];
                                const uint32_t t = lo ? lo : (uint32_t)src[t0 + r];
                                tok[k] = t;
                                if ((k & (XHI_SEG_SYMS - 1)) == 0) segpos[k / XHI_SEG_SYMS] = t0 + (uint32_t)r;
                                atomicAdd(&sm.hist[xhe_sym_of(t)], 1u);
                            }
                        }
                        __syncthreads();
                        if (tid == 0) sm.lz.ntok = tb + cnt;
                        __syncthreads();
                    }
                } else {    // literal-only fallback (xh_compress_no_matching :155-180)
                    for (int i = tid; i < XHE_SYMS; i += XHE_THREADS) sm.hist[i] = 0;
                    __syncthreads();
                    for (uint32_t i = tid; i < n; i += XHE_THREADS) { const uint32_t b = src[i]; tok[i] = b; atomicAdd(&sm.hist[b], 1u); }
                    for (uint32_t i = tid; i * XHI_SEG_SYMS < n; i += XHE_THREADS) segpos[i] = i * XHI_SEG_SYMS;
                    if (tid == 0) sm.lz.ntok = n;
                    __syncthreads();
                }
                if (tid == 0 && last) { tok[sm.lz.ntok] = XHE_TOK_EOS; sm.lz.ntok += 1; sm.hist[256] += 1; }
                __syncthreads();

                // ---- codes + exact size (xh_calc_compressed_len :181-188) ----
                if (EXACT && !literal_only) xhe_build_codes_reference(sm); else xhe_build_codes(sm);
                {
                    unsigned long long local = 0;      // bits << 32 | raw bytes
                    for (int s = tid; s < XHE_SYMS; s += XHE_THREADS) {
                        const uint32_t h = sm.hist[s];
                        if (h) {
                            const uint32_t ob = s >= 256 ? (uint32_t)((s >> 4) & 15) : 0u;
                            local += ((unsigned long long)(h * (sm.len()[s] + ob))) << 32;
                        }
                    }
                    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(FULL, local, d);
                    if (lane == 0) sm.wsum[warp] = local;
                    __syncthreads();
                    if (tid == 0) { unsigned long long t = 0; for (int i = 0; i < XHE_THREADS / 32; ++i) t += sm.wsum[i]; sm.bits_total = (uint32_t)(t >> 32); }
                    __syncthreads();
                }
                // raw bytes: counted from the tokens (only length symbols with nibble 15 carry them)
                {
                    uint32_t r = 0;
                    const uint32_t nt = sm.lz.ntok;
                    if (!literal_only) for (uint32_t k = tid; k < nt; k += XHE_THREADS) r += xhe_raw_of(tok[k]);
                    for (int d = 16; d; d >>= 1) r += __shfl_xor_sync(FULL, r, d);
                    if (lane == 0) sm.scan_tmp[warp] = r;
                    __syncthreads();
                    if (tid == 0) { uint32_t t = 0; for (int i = 0; i < XHE_THREADS / 32; ++i) t += sm.scan_tmp[i]; sm.raw_total = t; }
                    __syncthreads();
                }
                const uint32_t Btot = sm.bits_total;
                const uint32_t comp_len = 2 * (((Btot + 15) >> 4) + 1) + sm.raw_total;
                if (!literal_only && comp_len > n + (last ? 36u : 2u)) { literal_only = true; continue; }   // :274, :310
                break;
            }

            const uint32_t ntok = sm.lz.ntok;
            const uint32_t nsym = ntok - (last ? 1u : 0u);                // symbols that produce output (the index does not count the end-of-stream code)
            const uint32_t Btot = sm.bits_total;
            const uint32_t nwords = xhe_reserved(Btot);
            const uint32_t comp_len = 2 * nwords + sm.raw_total;
            if ((uint64_t)opos + 256 + comp_len > cap || nwords + 2 > XHE_WORD_CAP) { st = ST_BUF; break; }
            uint8_t* o = obase + opos + 256;
            // 256-byte table: symbol 2i in the low nibble of byte i (:284, :320)
            for (int i = tid; i < 256; i += XHE_THREADS) obase[opos + i] = (uint8_t)(sm.len()[2 * i] | (sm.len()[2 * i + 1] << 4));
            // ---- pass A: bit addresses, raw bytes, word addresses ----
            for (uint32_t i = tid; i < (nwords + 1) / 2 + 1; i += XHE_THREADS) bits32[i] = 0;
            if (tid == 0) { sm.bit_carry = 0; wpos[0] = 0; wpos[1] = 2; }
            __syncthreads();
            for (uint32_t k0 = 0; k0 < ntok; k0 += XHE_THREADS) {
                const uint32_t k = k0 + tid;
                uint32_t t = 0, sym = 0, cl = 0, ob = 0, raw = 0;
                if (k < ntok) {
                    t = tok[k]; sym = xhe_sym_of(t); cl = sm.len()[sym];
                    if (xhe_is_match(t)) { ob = (sym >> 4) & 15; raw = xhe_raw_of(t); }
                }
                unsigned long long v = ((unsigned long long)(cl + ob) << 32) | raw, incl = v;
                for (int d = 1; d < 32; d <<= 1) { const unsigned long long x = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += x; }
                if (lane == 31) sm.wsum[warp] = incl;
                __syncthreads();
                unsigned long long base = sm.bit_carry;
                for (int w2 = 0; w2 < warp; ++w2) base += sm.wsum[w2];
                const unsigned long long excl = base + incl - v;
                if (k < ntok) {
                    const uint32_t B0 = (uint32_t)(excl >> 32), R0 = (uint32_t)excl;
                    const uint32_t B1 = B0 + cl, B2 = B1 + ob;
                    if (ix && (k & (XHI_SEG_SYMS - 1)) == 0 && k < nsym) {   // the reader's state before symbol k (mask: see below)
                        const uint32_t F = xhe_reserved(B0), e = nent + k / XHI_SEG_SYMS;
                        const uint32_t left = nsym - k;
                        if (e < ix_room)
                            ix[e] = XhSeg{B0, opos + 256u + 2u * F + R0, c0 + segpos[k / XHI_SEG_SYMS],
                                          (16u * F - B0) | (k == 0 ? XHI_NEWCHUNK : 0u) | ((left < XHI_SEG_SYMS ? left : XHI_SEG_SYMS) << 8)};
                    }
                    xhe_put_bits(bits32, B0, sm.code()[sym], cl);
                    if (xhe_reserved(B1) > xhe_reserved(B0)) { const uint32_t j = xhe_reserved(B0); wpos[j] = 4 + 2 * (j - 2) + R0; }
                    if (xhe_is_match(t)) {
                        const uint32_t off = t & 0xFFFFu, lm3 = (t >> 16) + 1 - 3;
                        if (raw) {
                            uint8_t* rp = o + 4 + 2 * (B1 ? ((B1 + 15) >> 4) - 1 : 0) + R0;
                            if (raw == 1) rp[0] = (uint8_t)(lm3 - 15);
                            else { rp[0] = 0xFF; rp[1] = (uint8_t)lm3; rp[2] = (uint8_t)(lm3 >> 8); }
                        }
                        xhe_put_bits(bits32, B1, off - (1u << ob), ob);
                        if (xhe_reserved(B2) > xhe_reserved(B1)) { const uint32_t j = xhe_reserved(B1); wpos[j] = 4 + 2 * (j - 2) + R0 + raw; }
                    }
                }
                __syncthreads();
                if (tid == XHE_THREADS - 1) sm.bit_carry = excl + v;
                __syncthreads();
            }
            // ---- pass B: words to their addresses ----
            for (uint32_t j = tid; j < nwords; j += XHE_THREADS) {
                const uint32_t wv = bits16[j], p = wpos[j];
                o[p] = (uint8_t)wv; o[p + 1] = (uint8_t)(wv >> 8);
            }
            // index entries of this chunk: the bit window = the next `valid` bits of the bit string from the symbol's bit address on
            if (ix) {
                const uint32_t ne = (nsym + XHI_SEG_SYMS - 1) / XHI_SEG_SYMS;
                for (uint32_t s2 = tid; s2 < ne; s2 += XHE_THREADS) {
                    const uint32_t e = nent + s2;
                    if (e < ix_room) {
                        const uint32_t B0 = ix[e].mask, valid = ix[e].info & 63u;
                        const uint32_t j = B0 >> 4, r = B0 & 15u;
                        const uint32_t w0 = j < nwords ? bits16[j] : 0u, w1 = j + 1 < nwords ? bits16[j + 1] : 0u;
                        ix[e].mask = (((w0 << 16) | w1) << r) & (valid >= 32u ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> valid));
                    }
                }
                nent += ne;
            }
            __syncthreads();
            opos += 256 + comp_len;
        }
        if (tid == 0) { out_len[u] = (st == ST_OK) ? opos : 0; status[u] = st; if (nseg) nseg[u] = (st == ST_OK) ? nent : 0; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(XHE_THREADS, 4)
xhuff_encode_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                    const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                    const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                    uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint8_t* __restrict__ scratch, int eos_mode,
                    XhSeg* __restrict__ index, const uint64_t* __restrict__ index_off, uint32_t* __restrict__ nseg)
{
    xhuff_encode_body<false>(in, out, count, in_off, in_len, out_off, out_cap, out_len, status, scratch, eos_mode, index, index_off, nseg, XhPre{});
}

__global__ void __launch_bounds__(XHE_THREADS, 5)
xhuff_encode_exact_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t count,
                          const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                          const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
                          uint32_t* __restrict__ out_len, int32_t* __restrict__ status, uint8_t* __restrict__ scratch, int eos_mode,
                          XhSeg* __restrict__ index, const uint64_t* __restrict__ index_off, uint32_t* __restrict__ nseg, XhPre pre)
{
    xhuff_encode_body<true>(in, out, count, in_off, in_len, out_off, out_cap, out_len, status, scratch, eos_mode, index, index_off, nseg, pre);
}

// The one-unit description of a stream whose chunks are the units of a batch (eos_mode 1), for lz_prev_kernel / lz_find_kernel:
// desc_off[0] = in_off[0], desc_len[0] = bytes from there to the end of the last unit, seg_first = {0, chunks}; clears the work counters.
__global__ void xhuff_stream_desc_kernel(uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                                         uint64_t* __restrict__ desc_off, uint32_t* __restrict__ desc_len, uint32_t* __restrict__ seg_first,
                                         uint32_t* __restrict__ work_counter)
{
    if (threadIdx.x == 0) {
        const uint64_t total = in_off[count - 1] + in_len[count - 1] - in_off[0];
        desc_off[0] = in_off[0];
        desc_len[0] = (uint32_t)total;
        seg_first[0] = 0; seg_first[1] = (uint32_t)((total + XHE_CHUNK - 1) / XHE_CHUNK);
    }
    if (threadIdx.x < 4) work_counter[threadIdx.x] = 0;
}

} // namespace msc
