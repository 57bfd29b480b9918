/*
 * oracle_xdict.h -- TEST INFRASTRUCTURE (see oracle.h).
 * Restatement of the reference's hash-chain matcher, include/mscomp/XpressDictionary.h, level 3
 * (NiceLength 48, MaxChain 11, :28-36), HashBits 15, HashShift 5 (:57-60).
 *
 * Positions are kept as signed indices into the whole input instead of pointers; XD_NONE plays the role
 * of the NULL table entry (:99-102), which always fails the `x >= xend` test of Find (:161).
 *
 * Known divergence from the compiled reference: the reference never initialises `window` (:64-69), so
 * when Find() runs on a position that Fill() has not reached yet (only possible after a match longer than
 * ChunkSize in plain Xpress, src/xpress_compress.cpp:269) it reads indeterminate memory.  Here such a slot
 * reads as XD_NONE.
 */
#ifndef ORACLE_XDICT_H
#define ORACLE_XDICT_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XD_NONE      (-((int64_t)1 << 40))
#define XD_HASH_SIZE 0x8000
#define XD_NICE_LEN  48u
#define XD_MAX_CHAIN 11u

typedef struct {
    const uint8_t* d;       /* whole input */
    int64_t n;              /* its length (end), end2 = n-2 */
    int64_t max_offset, chunk_size, window_mask;
    int64_t* table;         /* XD_HASH_SIZE heads */
    int64_t* window;        /* 2*chunk_size prev links */
} xdict;

static int xd_init(xdict* x, const uint8_t* d, size_t n, int64_t max_offset, int64_t chunk_size)
{
    x->d = d; x->n = (int64_t)n; x->max_offset = max_offset; x->chunk_size = chunk_size;
    x->window_mask = 2 * chunk_size - 1;
    x->table = (int64_t*)malloc(XD_HASH_SIZE * sizeof(int64_t));
    x->window = (int64_t*)malloc((size_t)(2 * chunk_size) * sizeof(int64_t));
    if (!x->table || !x->window) { free(x->table); free(x->window); return 0; }
    for (int64_t i = 0; i < XD_HASH_SIZE; ++i) x->table[i] = XD_NONE;
    for (int64_t i = 0; i < 2 * chunk_size; ++i) x->window[i] = XD_NONE;
    return 1;
}
static void xd_free(xdict* x) { free(x->table); free(x->window); }

static unsigned xd_hash_update(unsigned h, uint8_t c) { return ((h << 5) ^ c) & (XD_HASH_SIZE - 1); }

/* Fill :104-118 -- insert positions [p, min(p+ChunkSize, end2)); returns the new filled-to position */
static int64_t xd_fill(xdict* x, int64_t p)
{
    const int64_t end2 = x->n - 2;
    if (p >= end2) return end2;
    int64_t pos = p & x->window_mask;
    const int64_t endx = (p + x->chunk_size < end2) ? p + x->chunk_size : end2;
    unsigned hash = xd_hash_update(x->d[p], x->d[p + 1]);
    while (p < endx) {
        hash = xd_hash_update(hash, x->d[p + 2]);
        x->window[pos++] = x->table[hash];
        x->table[hash] = p++;
    }
    return endx;
}

/* GetMatchLength :72-94 -- common run of a<b, stopping at `end`; one short when it reaches `end` */
static uint32_t xd_match_len(const uint8_t* d, int64_t a, int64_t b, int64_t end)
{
    const int64_t b_start = b;
    uint8_t a0, b0;
    do { a0 = d[a++]; b0 = d[b++]; } while (b < end && a0 == b0);
    return (uint32_t)(b - b_start - 1);
}

/* Find :145-183 */
static uint32_t xd_find(const xdict* x, int64_t p, uint32_t* offset)
{
    const int64_t xend = p - x->max_offset;
    uint32_t len = 2, chain = XD_MAX_CHAIN;
    for (int64_t c = x->window[p & x->window_mask]; chain && c >= xend; c = x->window[c & x->window_mask], --chain) {
        if (x->d[c] == x->d[p] && x->d[c + 1] == x->d[p + 1]) {
            const uint32_t l = xd_match_len(x->d, c, p, x->n);
            if (l > len) {
                *offset = (uint32_t)(p - c);
                len = l;
                if (len >= XD_NICE_LEN) break;
            }
        }
    }
    return len;
}
#endif
