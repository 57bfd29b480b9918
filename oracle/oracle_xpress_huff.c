/*
 * oracle_xpress_huff.c -- TEST INFRASTRUCTURE (see oracle.h).  Plain-C restatement of the reference's
 * Xpress-Huffman codec.
 *
 * Follows (file:line in /root/reference):
 *   bit reader ............ include/mscomp/Bitstream.h:34-107 (16-bit LE words, MSB first, raw bytes at the read pointer)
 *   bit writer ............ include/mscomp/Bitstream.h:111-148
 *   canonical decoder ..... include/mscomp/HuffmanDecoder.h:42-91 (tables), :93-103 (DecodeSymbol)
 *   chunk decode .......... src/xpress_huff_decompress.cpp:39-135 (checked loop :79-126; the fast loop :49-76 is equivalent)
 *   stream decode ......... src/xpress_huff_decompress.cpp:136-162
 *   LZ77 + histogram ...... src/xpress_huff_compress.cpp:52-153, literal-only fallback :155-180
 *   size prediction ....... src/xpress_huff_compress.cpp:181-194
 *   bit packing ........... src/xpress_huff_compress.cpp:195-245
 *   stream encode ......... src/xpress_huff_compress.cpp:247-331
 *   code construction ..... include/mscomp/HuffmanEncoder.h:58-127 (heap Huffman + rescale), :129-226 (package-merge),
 *                           include/mscomp/sorting.h:27-66 (stable sorts)
 *   matcher ............... include/mscomp/XpressDictionary.h via oracle_xdict.h, MaxOffset 0xFFFF, ChunkSize 0x10000
 */
#include "oracle.h"
#include "oracle_xdict.h"

#define XH_CHUNK     0x10000u
#define XH_SYMS      512
#define XH_EOS       0x100
#define XH_MAXBITS   15
#define XH_TABLEBITS 9
#define XH_INVALID   0xFFFFu
#define XH_MIN_DATA  (256 + 4)

static uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t rd32(const uint8_t* p) { return rd16(p) | (rd16(p + 2) << 16); }
static void wr16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void wr32(uint8_t* p, uint32_t v) { wr16(p, v); wr16(p + 2, v >> 16); }

size_t orc_xpress_huff_max_compressed_size(size_t in_len)
{   /* src/xpress_huff_compress.cpp:46 */
    return in_len + 34 + (256 + 2) + (256 + 2) * (in_len / XH_CHUNK);
}

/* ---- input bitstream (Bitstream.h:34-107) --------------------------------------------------------- */
typedef struct { const uint8_t* in; const uint8_t* end; uint32_t mask; unsigned bits; } ibits;

static void ib_init(ibits* b, const uint8_t* in, const uint8_t* end)
{ b->in = in + 4; b->end = end; b->mask = (rd16(in) << 16) | rd16(in + 2); b->bits = 32; }
static uint32_t ib_peek(const ibits* b, unsigned n) { return (b->mask >> 16) >> (16 - n); }
static int ib_mask_is_zero(const ibits* b) { return b->bits == 0 || (b->mask >> (32 - b->bits)) == 0; }
static size_t ib_remaining(const ibits* b) { return (size_t)(b->end - b->in); }
static void ib_skip(ibits* b, unsigned n)
{
    b->mask <<= n; b->bits -= n;
    if (b->bits < 16 && b->in + 2 <= b->end) {
        b->mask |= rd16(b->in) << (16 - b->bits);
        b->bits |= 0x10;
        b->in += 2;
    }
}

/* ---- canonical decoder (HuffmanDecoder.h) --------------------------------------------------------- */
typedef struct {
    uint32_t lims[XH_MAXBITS + 1], poss[XH_MAXBITS + 1];
    uint16_t syms[XH_SYMS];
    uint8_t lens[1 << XH_TABLEBITS];
} hdec;

static int hd_set_lengths(hdec* h, const uint8_t cl[XH_SYMS])
{
    uint32_t cnts[XH_MAXBITS + 1] = {0}, last = 0, index = 0;
    const uint32_t max_value = 1u << XH_MAXBITS;
    memset(h->syms, 0xFF, sizeof h->syms);
    for (int s = 0; s < XH_SYMS; ++s) ++cnts[cl[s]];
    cnts[0] = 0;
    h->lims[0] = 0; h->lims[XH_MAXBITS] = max_value;
    for (unsigned len = 1; len <= XH_TABLEBITS; ++len) {
        uint32_t inc = cnts[len] << (XH_MAXBITS - len);
        if (last + inc > max_value) return 0;
        h->lims[len] = (last += inc);
        uint32_t limit = last >> (XH_MAXBITS - XH_TABLEBITS);
        memset(h->lens + index, (int)len, limit - index); index = limit;
    }
    for (unsigned len = XH_TABLEBITS + 1; len < XH_MAXBITS; ++len) {
        uint32_t inc = cnts[len] << (XH_MAXBITS - len);
        if (last + inc > max_value) return 0;
        h->lims[len] = (last += inc);
    }
    if (last + cnts[XH_MAXBITS] > max_value) return 0;
    h->poss[0] = 0;
    for (unsigned len = 1; len <= XH_MAXBITS; ++len) h->poss[len] = h->poss[len - 1] + cnts[len - 1];
    uint32_t fill[XH_MAXBITS + 1];
    memcpy(fill, h->poss, sizeof fill);
    for (int s = 0; s < XH_SYMS; ++s) if (cl[s]) h->syms[fill[cl[s]]++] = (uint16_t)s;
    return 1;
}

/* DecodeSymbol :93-103 (bounds-aware variant) */
static unsigned hd_decode(const hdec* h, ibits* b)
{
    unsigned n, r = b->bits;
    const uint32_t x = (r < XH_MAXBITS) ? (ib_peek(b, r) << (XH_MAXBITS - r)) : ib_peek(b, XH_MAXBITS);
    if (x < h->lims[XH_TABLEBITS]) n = h->lens[x >> (XH_MAXBITS - XH_TABLEBITS)];
    else for (n = XH_TABLEBITS + 1; x >= h->lims[n]; ++n) ;
    if (n > r) return XH_INVALID;
    ib_skip(b, n);
    uint32_t s = h->poss[n] + ((x - h->lims[n - 1]) >> (XH_MAXBITS - n));
    return s >= XH_SYMS ? XH_INVALID : h->syms[s];
}

/* xpress_huff_decompress_chunk :39-135.  Returns 1 for end-of-stream, 0 for end-of-chunk, <0 for errors. */
static int xh_decode_chunk(const uint8_t** in_p, const uint8_t* in_end, uint8_t** out_p, const uint8_t* out_end,
                           const uint8_t* out_origin, const hdec* dec)
{
    ibits b;
    uint8_t* out = *out_p;
    const uint8_t* out_end_chunk = out + XH_CHUNK;
    ib_init(&b, *in_p, in_end);
    while (out < out_end_chunk || !ib_mask_is_zero(&b)) {                                   /* :79 */
        unsigned sym = hd_decode(dec, &b);
        if (sym == XH_INVALID) return ORC_DATA_ERROR;
        if (sym == XH_EOS && ib_remaining(&b) == 0 && ib_mask_is_zero(&b)) { *in_p = b.in; *out_p = out; return 1; }   /* :83 */
        if (sym < 0x100) {
            if (out == out_end) return ORC_BUF_ERROR;
            *out++ = (uint8_t)sym;
        } else {
            uint32_t len, off;
            if ((len = sym & 0xF) == 0xF) {
                if (ib_remaining(&b) < 1) return ORC_DATA_ERROR;
                if ((len = *b.in++) == 0xFF) {
                    if (ib_remaining(&b) < 2) return ORC_DATA_ERROR;
                    len = rd16(b.in); b.in += 2;
                    if (len == 0) {
                        if (ib_remaining(&b) < 4) return ORC_DATA_ERROR;
                        len = rd32(b.in); b.in += 4;
                    }
                    if (len < 0xF) return ORC_DATA_ERROR;
                    len -= 0xF;
                }
                len += 0xF;
            }
            len += 3;
            {
                const unsigned off_bits = (sym >> 4) & 0xF;
                if (off_bits > b.bits) return ORC_DATA_ERROR;
                off = ib_peek(&b, off_bits) + (1u << off_bits);
                ib_skip(&b, off_bits);
            }
            if ((size_t)(out - out_origin) < off) return ORC_DATA_ERROR;                    /* :113 */
            if ((size_t)(out_end - out) < len) return ORC_BUF_ERROR;                        /* :114 */
            for (uint32_t i = 0; i < len; ++i, ++out) *out = *(out - off);
        }
    }
    *out_p = out;
    *in_p = b.in;
    if (hd_decode(dec, &b) == XH_EOS && ib_remaining(&b) == 0 && ib_mask_is_zero(&b)) { *in_p = b.in; return 1; }   /* :129 */
    return 0;
}

int orc_xpress_huff_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len)
{
    const uint8_t* in_end = in + in_len;
    uint8_t* const out_start = out;
    const uint8_t* out_end = out + *out_len;
    hdec dec;
    uint8_t cl[XH_SYMS];
    int status;
    do {
        if ((size_t)(in_end - in) < XH_MIN_DATA) {                                          /* :145-149 */
            if (in != in_end) return ORC_DATA_ERROR;
            break;
        }
        for (int i = 0, i2 = 0; i < 256; ++i) { cl[i2++] = in[i] & 0xF; cl[i2++] = in[i] >> 4; }
        in += 256;
        if (!hd_set_lengths(&dec, cl)) return ORC_DATA_ERROR;
        status = xh_decode_chunk(&in, in_end, &out, out_end, out_start, &dec);
        if (status < 0) return status;
    } while (status != 1);
    *out_len = (size_t)(out - out_start);
    return ORC_OK;
}

/* ---- encoder ---------------------------------------------------------------------------------------- */

static unsigned ilog2_16(unsigned v) { unsigned r = 0; while (v >>= 1) ++r; return r; }

/* xh_compress_lz77 :52-153.  Writes the intermediate token stream into buf, fills counts[512]. */
static size_t xh_lz77(const uint8_t* base, size_t pos, int32_t in_len, size_t total, uint8_t* buf, uint32_t counts[XH_SYMS], xdict* d)
{
    int32_t rem = in_len;
    size_t ip = pos;
    uint8_t* out = buf;
    uint8_t* mask_out = NULL;
    uint32_t mask = 0;
    unsigned i = 0;

    xd_fill(d, (int64_t)pos);
    memset(counts, 0, XH_SYMS * sizeof(uint32_t));
    while (rem > 0) {
        mask = 0;
        mask_out = out; out += 4;
        for (i = 0; i < 32 && rem > 0; ++i) {
            uint32_t len, off = 0;
            mask >>= 1;
            if (rem >= 3 && (len = xd_find(d, (int64_t)ip, &off)) >= 3) {
                if (len > (uint32_t)rem) len = (uint32_t)rem;                               /* :93 */
                ip += len; rem -= (int32_t)len;
                len -= 3;
                mask |= 0x80000000u;
                const unsigned off_bits = ilog2_16((off | 1) & 0xFFFF);
                const uint8_t sym = (uint8_t)((off_bits << 4) | (len < 0xF ? len : 0xF));
                ++counts[0x100 | sym];
                off ^= 1u << off_bits;
                out[0] = sym; wr16(out + 1, off); out += 3;
                if (len > 0xFFFF) { out[0] = 0xFF; wr16(out + 1, 0); wr32(out + 3, len); out += 7; }
                else if (len >= 0xFF + 0xF) { out[0] = 0xFF; wr16(out + 1, len); out += 3; }
                else if (len >= 0xF) { *out++ = (uint8_t)(len - 0xF); }
            } else {
                ++counts[*out++ = base[ip++]];
                --rem;
            }
        }
        wr32(mask_out, mask);
    }
    mask >>= (32 - i);                                                                       /* :125 */
    if (pos + (size_t)in_len == total) {                                                     /* :126-144 EOS */
        if (i == 32) { wr32(out, 1); out += 4; }
        else mask |= 1u << i;
        wr32(out, 0); out += 3;
        ++counts[XH_EOS];
    }
    if (mask_out) wr32(mask_out, mask);
    return (size_t)(out - buf);
}

/* xh_compress_no_matching :155-180 */
static size_t xh_no_matching(const uint8_t* in, size_t in_len, int is_end, uint8_t* buf, uint32_t counts[XH_SYMS])
{
    const uint8_t* in_end = in + in_len;
    uint8_t* out = buf;
    memset(counts, 0, XH_SYMS * sizeof(uint32_t));
    while (in_end - in > 32) {
        wr32(out, 0); out += 4;
        memcpy(out, in, 32); out += 32;
        for (int i = 0; i < 32; ++i) ++counts[*in++];
    }
    const size_t rem = (size_t)(in_end - in);
    wr32(out, 0); out += 4;
    memcpy(out, in, rem); out += rem;
    while (in < in_end) ++counts[*in++];
    if (is_end) {
        if (rem == 32) { wr32(out, 1); out += 4; }
        else wr32(out - rem - 4, 1u << rem);
        wr32(out, 0); out += 3;
        ++counts[XH_EOS];
    }
    return (size_t)(out - buf);
}

/* xh_calc_compressed_len :181-188 */
static size_t xh_calc_len(const uint8_t lens[XH_SYMS], const uint32_t counts[XH_SYMS], size_t buf_len)
{
    size_t sym_bits = 16;
    uint32_t lit = 0, mat = 0;
    for (int i = 0; i < 0x100; ++i) { sym_bits += (size_t)lens[i] * counts[i]; lit += counts[i]; }
    for (int i = 0x100; i < XH_SYMS; ++i) { sym_bits += (size_t)(lens[i] + ((i >> 4) & 0xF)) * counts[i]; mat += counts[i]; }
    return (sym_bits + 15) / 16 * 2 + (buf_len - (lit + mat * 3 + (lit + mat + 31) / 32 * 4));
}
/* xh_calc_compressed_len_no_matching :189-194 */
static size_t xh_calc_len_nm(const uint8_t lens[XH_SYMS], const uint32_t counts[XH_SYMS])
{
    size_t sym_bits = 16;
    for (int i = 0; i <= 0x100; ++i) sym_bits += (size_t)lens[i] * counts[i];
    return (sym_bits + 15) / 16 * 2;
}

/* ---- HuffmanEncoder::CreateCodes :58-127 ---- */
typedef struct { uint16_t codes[XH_SYMS]; uint8_t lens[XH_SYMS]; } henc;

/* HEAP_PUSH / HEAP_POP, HuffmanEncoder.h:30-54: 1-based min-heap of node ids keyed by weights[], id 0 is a sentinel */
static void heap_push(unsigned* heap, unsigned* heap_len, const uint32_t* weights, unsigned x)
{
    unsigned j = ++*heap_len;
    heap[j] = x;
    while (weights[x] < weights[heap[j >> 1]]) { heap[j] = heap[j >> 1]; j >>= 1; }
    heap[j] = x;
}
static void heap_pop(unsigned* heap, unsigned* heap_len, const uint32_t* weights)
{
    unsigned i = 1, t = heap[1] = heap[(*heap_len)--];
    for (;;) {
        unsigned j = i << 1;
        if (j > *heap_len) break;
        if (j < *heap_len && weights[heap[j + 1]] < weights[heap[j]]) ++j;
        if (weights[t] < weights[heap[j]]) break;
        heap[i] = heap[j];
        i = j;
    }
    heap[i] = t;
}

static void he_create_codes(henc* e, const uint32_t counts[XH_SYMS])
{
    uint32_t weights[XH_SYMS * 2];
    memset(e->codes, 0, sizeof e->codes);
    weights[0] = 0;
    for (int i = 0; i < XH_SYMS; ++i) weights[i + 1] = (counts[i] == 0 ? 1 : counts[i]) << 8;
    for (;;) {
        unsigned heap_len = 0, heap[XH_SYMS + 2] = {0}, n_nodes = XH_SYMS, parents[XH_SYMS * 2];
        for (unsigned i = 1; i <= XH_SYMS; ++i) heap_push(heap, &heap_len, weights, i);
        memset(parents, 0, sizeof parents);
        while (heap_len > 1) {
            unsigned n1 = heap[1]; heap_pop(heap, &heap_len, weights);
            unsigned n2 = heap[1]; heap_pop(heap, &heap_len, weights);
            parents[n1] = parents[n2] = ++n_nodes;
            uint32_t d1 = weights[n1] & 0xFF, d2 = weights[n2] & 0xFF;
            weights[n_nodes] = ((weights[n1] & 0xFFFFFF00u) + (weights[n2] & 0xFFFFFF00u)) | (1 + (d1 > d2 ? d1 : d2));
            heap_push(heap, &heap_len, weights, n_nodes);
        }
        int too_long = 0;
        for (unsigned i = 1; i <= XH_SYMS; ++i) {
            unsigned j = 0, k = i;
            while (parents[k] > 0) { k = parents[k]; ++j; }
            e->lens[i - 1] = (uint8_t)j;
            if (j > XH_MAXBITS) too_long = 1;
        }
        if (!too_long) break;
        for (unsigned i = 1; i <= XH_SYMS; ++i) weights[i] = (1 + (weights[i] >> 9)) << 8;
    }
    unsigned mn = e->lens[0], mx = mn;
    for (int i = 1; i < XH_SYMS; ++i) { if (e->lens[i] > mx) mx = e->lens[i]; else if (e->lens[i] < mn) mn = e->lens[i]; }
    uint16_t code = 0;
    for (unsigned n = mn; n <= mx; ++n) {
        for (int i = 0; i < XH_SYMS; ++i) if (e->lens[i] == n) e->codes[i] = code++;
        code <<= 1;
    }
}

/* stable sort of syms by key[syms[x]] -- sorting.h:27-66 (any stable sort gives the same order) */
static void stable_sort_u32(uint16_t* syms, unsigned len, const uint32_t* key)
{
    for (unsigned i = 1; i < len; ++i) {
        uint16_t x = syms[i]; uint32_t c = key[x]; unsigned j = i;
        while (j > 0 && key[syms[j - 1]] > c) { syms[j] = syms[j - 1]; --j; }
        syms[j] = x;
    }
}
static void stable_sort_u8(uint16_t* syms, unsigned len, const uint8_t* key)
{
    for (unsigned i = 1; i < len; ++i) {
        uint16_t x = syms[i]; uint8_t c = key[x]; unsigned j = i;
        while (j > 0 && key[syms[j - 1]] > c) { syms[j] = syms[j - 1]; --j; }
        syms[j] = x;
    }
}

/* ---- HuffmanEncoder::CreateCodesSlow :129-226 (package-merge) ---- */
typedef struct { uint8_t symbols[XH_SYMS]; uint32_t count; } pm_col;

static int he_create_codes_slow(henc* e, const uint32_t counts[XH_SYMS])
{
    uint16_t by_count[XH_SYMS], by_len[XH_SYMS];
    unsigned len = 0;
    memset(e->codes, 0, sizeof e->codes);
    memset(e->lens, 0, sizeof e->lens);
    for (int i = 0; i < XH_SYMS; ++i) if (counts[i]) { by_count[len] = (uint16_t)i; by_len[len++] = (uint16_t)i; e->lens[i] = XH_MAXBITS; }
    stable_sort_u32(by_count, len, counts);
    if (len == 1) { e->lens[by_count[0]] = 1; return 1; }
    pm_col* a = (pm_col*)malloc(XH_SYMS * sizeof(pm_col));
    pm_col* b = (pm_col*)malloc(XH_SYMS * sizeof(pm_col));
    if (!a || !b) { free(a); free(b); return 0; }
    pm_col *cols = a, *next_cols = b;
    unsigned cols_len = 0, next_len = 0;
    for (unsigned j = 0; j < XH_MAXBITS; ++j) {
        unsigned cp = 0, pos = 0;
        while ((cols_len - cp + len - pos) > 1) {
            memset(next_cols + next_len, 0, sizeof(pm_col));
            for (int i = 0; i < 2; ++i) {
                if (pos >= len || (cp < cols_len && cols[cp].count < counts[by_count[pos]])) {
                    next_cols[next_len].count += cols[cp].count;
                    for (int s = 0; s < XH_SYMS; ++s) next_cols[next_len].symbols[s] += cols[cp].symbols[s];
                    ++cp;
                } else {
                    next_cols[next_len].count += counts[by_count[pos]];
                    ++next_cols[next_len].symbols[by_count[pos]];
                    ++pos;
                }
            }
            ++next_len;
        }
        if (cp < cols_len) { for (int i = 0; i < XH_SYMS; ++i) e->lens[i] -= cols[cp].symbols[i]; }
        else if (pos < len) { --e->lens[by_count[pos]]; }
        pm_col* t = cols; cols = next_cols; next_cols = t;
        cols_len = next_len; next_len = 0;
    }
    free(a); free(b);
    stable_sort_u8(by_len, len, e->lens);
    for (unsigned i = 1; i < len; ++i)
        e->codes[by_len[i]] = (uint16_t)((e->codes[by_len[i - 1]] + 1) << (e->lens[by_len[i]] - e->lens[by_len[i - 1]]));
    return 1;
}

/* ---- output bitstream (Bitstream.h:111-148) ---- */
typedef struct { uint8_t* out; uint8_t* p0; uint8_t* p1; uint32_t mask; unsigned bits; } obits;
static void ob_init(obits* b, uint8_t* out) { b->out = out + 4; b->p0 = out; b->p1 = out + 2; b->mask = 0; b->bits = 0; }
static void ob_write(obits* b, uint32_t v, unsigned n)
{
    b->bits += n;
    if (n) b->mask |= v << (32 - b->bits);        /* n == 0 only with v == 0 (offset 1); avoids the shift by 32 */
    if (b->bits > 16) {
        wr16(b->p0, b->mask >> 16);
        b->mask <<= 16; b->bits &= 0xF;
        b->p0 = b->p1; b->p1 = b->out; b->out += 2;
    }
}
static void ob_finish(obits* b) { wr16(b->p0, b->mask >> 16); wr16(b->p1, 0); }

/* xh_compress_encode :195-245 */
static void xh_encode(const uint8_t* in, const uint8_t* in_end, uint8_t* out, const henc* e)
{
    obits b;
    ob_init(&b, out);
    while (in < in_end) {
        unsigned i;
        uint32_t mask;
        for (i = 32, mask = rd32(in), in += 4; mask && in < in_end; --i, mask >>= 1) {
            if (mask & 1) {
                const uint8_t sym = *in++;
                const uint32_t off = rd16(in); in += 2;
                ob_write(&b, e->codes[0x100 | sym], e->lens[0x100 | sym]);
                if ((sym & 0xF) == 0xF) {
                    const uint8_t len8 = *in++;
                    *b.out++ = len8;
                    if (len8 == 0xFF) {
                        const uint32_t len16 = rd16(in); in += 2;
                        wr16(b.out, len16); b.out += 2;
                        if (len16 == 0) { wr32(b.out, rd32(in)); b.out += 4; in += 4; }
                    }
                }
                ob_write(&b, off, sym >> 4);
            } else {
                ob_write(&b, e->codes[*in], e->lens[*in]); ++in;
            }
        }
        const uint8_t* end = (in + i < in_end) ? in + i : in_end;
        for (; in != end; ++in) ob_write(&b, e->codes[*in], e->lens[*in]);
    }
    ob_finish(&b);
}

int orc_xpress_huff_compress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len_p)
{
    if (in_len == 0) { *out_len_p = 0; return ORC_OK; }                                      /* :249 */
    uint8_t* buf = (uint8_t*)malloc(in_len >= XH_CHUNK ? 0x1200C : ((in_len + 31) / 32 * 36 + 4 + 8));
    henc* e = (henc*)malloc(sizeof(henc));
    xdict d;
    if (!buf || !e || !xd_init(&d, in, in_len, 0xFFFF, XH_CHUNK)) { free(buf); free(e); return ORC_MEM_ERROR; }

    uint8_t* const out_orig = out;
    size_t out_len = *out_len_p, pos = 0, left = in_len;
    uint32_t counts[XH_SYMS];
    int rc = ORC_OK;
    while (left > 0) {
        const int last = left <= XH_CHUNK;                                                  /* :262 / :298 */
        const size_t n = last ? left : XH_CHUNK;
        size_t buf_len = xh_lz77(in, pos, (int32_t)n, in_len, buf, counts, &d);
        he_create_codes(e, counts);
        size_t comp_len = xh_calc_len(e->lens, counts, buf_len);
        if (comp_len > (last ? n + 36 : XH_CHUNK + 2)) {                                     /* :274 / :310 */
            buf_len = xh_no_matching(in + pos, n, last, buf, counts);
            if (!he_create_codes_slow(e, counts)) { rc = ORC_MEM_ERROR; break; }
            comp_len = xh_calc_len_nm(e->lens, counts);
        }
        if (out_len < 256 + comp_len) { rc = ORC_BUF_ERROR; break; }
        for (int i = 0; i < XH_SYMS; i += 2) *out++ = (uint8_t)(e->lens[i] | (e->lens[i + 1] << 4));
        xh_encode(buf, buf + buf_len, out, e);
        out += comp_len; out_len -= 256 + comp_len;
        pos += n; left -= n;
    }
    xd_free(&d); free(buf); free(e);
    if (rc != ORC_OK) return rc;
    *out_len_p = (size_t)(out - out_orig);
    return ORC_OK;
}
