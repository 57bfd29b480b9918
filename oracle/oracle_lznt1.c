/*
 * oracle_lznt1.c -- TEST INFRASTRUCTURE (see oracle.h).  Plain-C restatement of the reference's LZNT1.
 *
 * Follows (file:line in /root/reference):
 *   chunk decode ......... src/lznt1_decompress.cpp:37-121   (bounds-checked loop :77-116 semantics)
 *   chunk header / driver  src/lznt1_decompress.cpp:122-209, :230-273 and the one-shot wrapper
 *                          include/mscomp/internal.h:616-630
 *   chunk encode ......... src/lznt1_compress.cpp:49-94
 *   one-shot encode ...... src/lznt1_compress.cpp:233-273
 *   matcher .............. include/mscomp/LZNT1Dictionary.h:93-106 (Fill), :114-143 (Find)
 *
 * Deliberate deviation (undefined behaviour upstream, never tested): a literal that would be written past
 * byte 4096 of a chunk is reported as ORC_DATA_ERROR (the reference's checked loop has that bounds check
 * commented out at src/lznt1_decompress.cpp:111 and writes out of bounds).
 */
#include "oracle.h"
#include <string.h>
#include <stdlib.h>

#define CHUNK 4096u

static unsigned rd16(const uint8_t* p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }

size_t orc_lznt1_max_compressed_size(size_t in_len)
{   /* src/lznt1_compress.cpp:27 */
    return in_len + 3 + 2 * ((in_len + CHUNK - 1) / CHUNK);
}

int orc_lznt1_decompress_chunk(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len)
{
    size_t ip = 0, op = 0, pow2 = 16;
    unsigned mask = 0xFFF, shift = 12;
    while (ip < n) {
        unsigned flags = in[ip++];
        for (int bit = 0; bit < 8; ++bit) {
            if (ip == n) { *out_len = op; return ORC_OK; }           /* :86 */
            if (flags & (1u << bit)) {
                if (ip + 2 > n) return ORC_DATA_ERROR;                   /* :90 */
                while (op > pow2) { pow2 <<= 1; mask >>= 1; --shift; }   /* :91 */
                unsigned sym = rd16(in + ip); ip += 2;
                size_t off = (sym >> shift) + 1, len = (sym & mask) + 3;
                if (off > op) return ORC_DATA_ERROR;                     /* :98 */
                if (op + len > cap) return ORC_DATA_ERROR;               /* :99 (mapped to DATA_ERROR at :161/:176) */
                for (size_t i = 0; i < len; ++i, ++op) out[op] = out[op - off];
            } else {
                if (op == cap) return ORC_DATA_ERROR;                    /* deviation, see header */
                out[op++] = in[ip++];
            }
        }
    }
    *out_len = op;
    return ORC_OK;
}

int orc_lznt1_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len)
{
    const size_t cap = *out_len;
    size_t ip = 0, op = 0;
    int leftover = 0;                      /* decoded bytes that did not fit (state->out_avail != 0) */
    uint8_t tmp[CHUNK];
    while (cap - op > 0 && in_len - ip >= 2) {                           /* :254 */
        const unsigned h = rd16(in + ip);
        if (h == 0) {                                                    /* :128-134 end-of-stream marker */
            if (in_len - ip != 2) return ORC_DATA_ERROR;
            *out_len = op;
            return ORC_OK;
        }
        const size_t in_size = (h & 0x0FFF) + 3;
        if (in_size > in_len - ip) return ORC_BUF_ERROR;                 /* :136-143 partial chunk -> wrapper :627 */
        if ((h & 0x7000) != 0x3000) return ORC_DATA_ERROR;               /* :152 */
        if (h & 0x8000) {
            size_t got;
            if (cap - op < CHUNK) {                                      /* :155-172 bounce buffer */
                if (orc_lznt1_decompress_chunk(in + ip + 2, in_size - 2, tmp, CHUNK, &got) != ORC_OK) return ORC_DATA_ERROR;
                size_t copy = got < cap - op ? got : cap - op;
                memcpy(out + op, tmp, copy);
                op += copy;
                if (copy < got) leftover = 1;
            } else {                                                     /* :173-186 */
                if (orc_lznt1_decompress_chunk(in + ip + 2, in_size - 2, out + op, CHUNK, &got) != ORC_OK) return ORC_DATA_ERROR;
                op += got;
            }
        } else {                                                         /* :188-206 stored chunk */
            const size_t sz = in_size - 2;
            if (cap - op < sz) { memcpy(out + op, in + ip + 2, cap - op); op = cap; leftover = 1; }
            else { memcpy(out + op, in + ip + 2, sz); op += sz; }
        }
        ip += in_size;
    }
    /* lznt1_is_possible_stream_end :224-229 + wrapper internal.h:627 */
    const size_t rem = in_len - ip;
    if (!leftover && (rem == 0 || (rem == 1 && in[ip] == 0))) { *out_len = op; return ORC_OK; }
    return ORC_BUF_ERROR;
}

/* ---- encoder ------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t head[65536];    /* first position (ascending list) per 2-byte prefix, -1 = empty */
    int32_t count[65536];
    int16_t next[CHUNK];    /* next (larger) position with the same prefix */
} lznt1_dict;

/* LZNT1Dictionary::Fill :93-106 -- positions 0..len-3 appended to the list of their 2-byte prefix */
static void dict_fill(lznt1_dict* d, const uint8_t* data, int len)
{
    memset(d->head, 0xFF, sizeof d->head);
    memset(d->count, 0, sizeof d->count);
    for (int p = len - 3; p >= 0; --p) {
        unsigned idx = ((unsigned)data[p] << 8) | data[p + 1];
        d->next[p] = (int16_t)d->head[idx];
        d->head[idx] = p;
        d->count[idx]++;
    }
}

/* LZNT1Dictionary::Find :114-143 -- longest match >= 3 among earlier positions with the same 2-byte prefix,
 * ascending, a later one wins only if strictly longer, stop at max_len; the list's newest entry is never
 * examined (size = sizes[idx]-1, :120). */
static int dict_find(const lznt1_dict* d, const uint8_t* data, int p, int max_len, int* offset)
{
    if (max_len >= 3 && p > 0) {
        unsigned idx = ((unsigned)data[p] << 8) | data[p + 1];
        const uint8_t z = data[p + 2];
        int size = d->count[idx] - 1, len = 0, found = -1;
        int j = d->head[idx];
        for (int k = 0; k < size && j < p; ++k, j = d->next[j]) {
            if (data[j + 2] == z) {
                int i = 3;
                while (i < max_len && data[p + i] == data[j + i]) ++i;
                if (i > len) { found = j; len = i; if (len == max_len) break; }
            }
        }
        if (len >= 3) { *offset = p - found; return len; }
    }
    return 0;
}

/* lznt1_compress_chunk :49-94.  Returns in_len when the chunk must be stored raw or does not fit. */
static unsigned compress_chunk(const uint8_t* in, unsigned in_len, uint8_t* out, size_t out_len, lznt1_dict* d)
{
    unsigned in_pos = 0, out_pos = 0, rem = in_len, pow2 = 0x10, mask3 = 0x1002, shift = 12;
    dict_fill(d, in, (int)in_len);
    while (out_pos < out_len && rem) {
        unsigned i = 0, pos = 0, bits = 0;
        uint8_t bytes[16];
        for (; i < 8 && out_pos < out_len && rem; ++i) {
            bits >>= 1;
            while (pow2 < in_pos) { pow2 <<= 1; mask3 = (mask3 >> 1) + 1; --shift; }      /* :66 */
            int off = 0, len = dict_find(d, in, (int)in_pos, (int)(rem < mask3 ? rem : mask3), &off);
            if (len > 0) {
                unsigned sym = (unsigned)(((off - 1) << shift) | (len - 3));
                bytes[pos++] = (uint8_t)sym; bytes[pos++] = (uint8_t)(sym >> 8);
                bits |= 0x80;
                in_pos += (unsigned)len; rem -= (unsigned)len;
            } else {
                bytes[pos++] = in[in_pos++];
                --rem;
            }
        }
        unsigned end = out_pos + 1 + pos;
        if (end >= in_len || end > out_len) return in_len;                                   /* :86 */
        out[out_pos] = (uint8_t)(bits >> (8 - i));
        memcpy(out + out_pos + 1, bytes, pos);
        out_pos += 1 + pos;
    }
    return rem ? in_len : out_pos;
}

int orc_lznt1_compress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len_p)
{   /* src/lznt1_compress.cpp:233-273; *out_len == 0 is undefined upstream (:239), rejected here */
    const size_t out_len = *out_len_p;
    size_t out_pos = 0, in_pos = 0;
    if (out_len == 0) return ORC_BUF_ERROR;
    lznt1_dict* d = (lznt1_dict*)malloc(sizeof *d);
    if (!d) return ORC_MEM_ERROR;
    while (out_pos < out_len - 1 && in_pos < in_len) {
        const unsigned in_size = (unsigned)(in_len - in_pos < CHUNK ? in_len - in_pos : CHUNK);
        unsigned out_size = compress_chunk(in + in_pos, in_size, out + out_pos + 2, out_len - out_pos - 2, d), flags;
        if (out_size < in_size) flags = 0xB000;
        else {
            if (out_pos + 2 + in_size > out_len) { free(d); return ORC_BUF_ERROR; }
            out_size = in_size; flags = 0x3000;
            memcpy(out + out_pos + 2, in + in_pos, out_size);
        }
        unsigned header = flags | (out_size - 1);
        out[out_pos] = (uint8_t)header; out[out_pos + 1] = (uint8_t)(header >> 8);
        out_pos += out_size + 2;
        in_pos += in_size;
    }
    free(d);
    if (in_pos < in_len) return ORC_BUF_ERROR;
    if (out_len - out_pos >= 2) out[out_pos] = out[out_pos + 1] = 0;     /* :268-271 terminator, not counted */
    *out_len_p = out_pos;
    return ORC_OK;
}
