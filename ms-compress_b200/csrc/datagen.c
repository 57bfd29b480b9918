/*
 * datagen.c -- deterministic synthetic byte streams for the benchmark and the tests (host side, plain C).
 *
 * Definitions are those of SURVEY.md section 8(d): PRNG = splitmix64 with state initialised to `seed`;
 *   text   : 1024-word vocabulary of 2..10 lower-case letters + ' ', words drawn with a skewed index
 *   random : the PRNG output as little-endian bytes (incompressible)
 *   runs   : byte runs of 1..1024 with every 8th draw a short repeated 8-byte pattern
 * Buffer i of a batch uses seed 42+i; a "1 GiB stream" is 1024 x 1 MiB blocks with seeds 42..1065.
 * The generator digests are pinned in tests/test_datagen.py against SURVEY.md Appendix B.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static uint64_t sm64(uint64_t* x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void msgen_text(uint8_t* out, size_t n, uint64_t seed)
{
    enum { V = 1024 };
    uint8_t words[V][12];
    uint8_t wlen[V];
    uint64_t s = seed;
    for (int w = 0; w < V; ++w) {
        uint64_t r = sm64(&s);
        int L = 2 + (int)(r % 9);
        for (int i = 0; i < L; ++i) { r = sm64(&s); words[w][i] = (uint8_t)(97 + r % 26); }
        words[w][L] = ' ';
        wlen[w] = (uint8_t)(L + 1);
    }
    size_t p = 0;
    while (p < n) {
        uint64_t r = sm64(&s);
        uint64_t a = r & 0xFFFF, b = (r >> 16) & 0xFFFF, c = (r >> 32) & 0xFFFF;
        uint64_t u = (((a * b) >> 16) * c) >> 16;
        unsigned idx = (unsigned)((u * V) >> 16);
        size_t l = wlen[idx];
        if (l > n - p) l = n - p;
        memcpy(out + p, words[idx], l);
        p += l;
    }
}

void msgen_random(uint8_t* out, size_t n, uint64_t seed)
{
    uint64_t s = seed;
    size_t p = 0;
    while (p < n) {
        uint64_t r = sm64(&s);
        for (int i = 0; i < 8 && p < n; ++i) out[p++] = (uint8_t)(r >> (8 * i));
    }
}

void msgen_runs(uint8_t* out, size_t n, uint64_t seed)
{
    uint64_t s = seed, k = 0;
    size_t p = 0;
    while (p < n) {
        uint64_t r = sm64(&s);
        ++k;
        if (k % 8 == 0) {
            size_t m = 1 + ((r >> 24) & 31);
            uint64_t q = sm64(&s);
            for (size_t i = 0; i < m && p < n; ++i) out[p++] = (uint8_t)(q >> (8 * (i & 7)));
        } else {
            size_t m = 1 + ((r >> 8) & 0x3FF);
            if (m > n - p) m = n - p;
            memset(out + p, (int)(r & 0xFF), m);
            p += m;
        }
    }
}

/* kind: 0 text, 1 random, 2 runs.  Fills `count` consecutive buffers of `each` bytes, buffer i with seed0+i. */
void msgen_batch(uint8_t* out, size_t each, size_t count, uint64_t seed0, int kind)
{
    for (size_t i = 0; i < count; ++i) {
        uint8_t* dst = out + i * each;
        if (kind == 0) msgen_text(dst, each, seed0 + i);
        else if (kind == 1) msgen_random(dst, each, seed0 + i);
        else msgen_runs(dst, each, seed0 + i);
    }
}
