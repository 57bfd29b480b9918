/*
 * oracle_xpress.c -- TEST INFRASTRUCTURE (see oracle.h).  Plain-C restatement of the reference's plain Xpress.
 *
 * Follows (file:line in /root/reference):
 *   decode ....... src/xpress_decompress.cpp:405-462 (bounds-checked loop :425-460; the unchecked fast loop
 *                  :119-198 computes the same thing), symbol reader :62-113
 *   encode ....... src/xpress_compress.cpp:240-347
 *   matcher ...... include/mscomp/XpressDictionary.h (see oracle_xdict.h), window 0x2000 (src/xpress_compress.cpp:28)
 */
#include "oracle.h"
#include "oracle_xdict.h"

static uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t rd32(const uint8_t* p) { return rd16(p) | (rd16(p + 2) << 16); }
static void wr16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void wr32(uint8_t* p, uint32_t v) { wr16(p, v); wr16(p + 2, v >> 16); }

size_t orc_xpress_max_compressed_size(size_t in_len) { return in_len + 4 + 4 * (in_len / 32); }  /* :30 */

/* src/xpress_decompress.cpp:41 -- all 1 bits above all 0 bits */
static int set_bits_are_highest(uint32_t x) { x = ~x; return !((x + 1) & x); }

int orc_xpress_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len)
{
    const size_t cap = *out_len;
    size_t ip = 0, op = 0;
    long half = -1;                                     /* index of the byte whose high nibble is still unused */
    if (in_len < 5) {                                   /* :414-418 (sic: a lone FFFFFFFF word is rejected) */
        if (in_len == 0 || (in_len == 4 && rd32(in) != 0xFFFFFFFFu)) { *out_len = 0; return ORC_OK; }
        return ORC_DATA_ERROR;
    }
    while (ip + 4 <= in_len) {
        uint32_t flags = rd32(in + ip), flagged = flags & 0x80000000u;
        flags = (flags << 1) | 1;
        ip += 4;
        do {
            if (ip == in_len) {                                                   /* :433-438 */
                if (!flagged || !set_bits_are_highest(flags)) return ORC_DATA_ERROR;
                *out_len = op;
                return ORC_OK;
            } else if (flagged) {                                                 /* _READ_SYMBOL :85-112 */
                if (ip + 2 > in_len) return ORC_DATA_ERROR;
                uint32_t sym = rd16(in + ip), off = (sym >> 3) + 1, len = sym & 7;
                ip += 2;
                if (len == 7) {
                    if (half >= 0) { len = in[half] >> 4; half = -1; }
                    else if (ip == in_len) return ORC_DATA_ERROR;
                    else { half = (long)ip; len = in[ip++] & 0xF; }
                    if (len == 0xF) {
                        if (ip == in_len) return ORC_DATA_ERROR;
                        if ((len = in[ip++]) == 0xFF) {
                            if (ip + 2 > in_len) return ORC_DATA_ERROR;
                            len = rd16(in + ip); ip += 2;
                            if (len == 0) {
                                if (ip + 4 > in_len) return ORC_DATA_ERROR;
                                len = rd32(in + ip); ip += 4;
                            }
                            if (len < 0xF + 0x7) return ORC_DATA_ERROR;
                            len -= 0xF + 0x7;
                        }
                        len += 0xF;
                    }
                    len += 0x7;
                }
                len += 0x3;
                if (off > op) return ORC_DATA_ERROR;                              /* :442 */
                if (op + (size_t)len > cap) return ORC_BUF_ERROR;                 /* :443 */
                for (uint32_t i = 0; i < len; ++i, ++op) out[op] = out[op - off];
            } else if (op == cap) return ORC_BUF_ERROR;                           /* :455 */
            else out[op++] = in[ip++];
            flagged = flags & 0x80000000u;
            flags <<= 1;
        } while (flags);
    }
    return ORC_DATA_ERROR;                                                        /* :461 */
}

int orc_xpress_compress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len_p)
{
    const size_t out_len = *out_len_p;
    size_t ip = 0, op = 0, flags_at = 0;
    int64_t filled_to = 0;
    uint32_t flags = 0;
    unsigned flag_count;
    long half = -1;
    xdict d;

    if (in_len == 0) {                                                            /* :253-259 */
        if (out_len < 4) return ORC_BUF_ERROR;
        wr32(out, 0xFFFFFFFFu); *out_len_p = 4; return ORC_OK;
    }
    if (out_len < 5) return ORC_BUF_ERROR;
    if (!xd_init(&d, in, in_len, 0x2000, 0x2000)) return ORC_MEM_ERROR;

#define FAIL_BUF do { xd_free(&d); return ORC_BUF_ERROR; } while (0)
    op = 4;
    out[op++] = in[ip++];                                                         /* :262-264 first byte literal */
    flag_count = 1;
    while (ip + 2 < in_len && op + 1 < out_len) {                                 /* :266 in < in_end2 && out < out_end1 */
        uint32_t len, off = 0;
        if (filled_to <= (int64_t)ip) filled_to = xd_fill(&d, filled_to);         /* :269 */
        flags <<= 1;
        if ((len = xd_find(&d, (int64_t)ip, &off)) < 3) { out[op++] = in[ip++]; }
        else {
            ip += len;
            len -= 3;
            wr16(out + op, ((off - 1) << 3) | (len < 7 ? len : 7));
            op += 2;
            if (len >= 7) {
                len -= 7;
                if (half >= 0) { out[half] |= (uint8_t)((len < 0xF ? len : 0xF) << 4); half = -1; }
                else {
                    if (op >= out_len) FAIL_BUF;
                    half = (long)op; out[op++] = (uint8_t)(len < 0xF ? len : 0xF);
                }
                if (len >= 0xF) {
                    len -= 0xF;
                    if (op >= out_len) FAIL_BUF;
                    out[op++] = (uint8_t)(len < 0xFF ? len : 0xFF);
                    if (len >= 0xFF) {
                        len += 0xF + 0x7;
                        if (len <= 0xFFFF) {
                            if (op + 2 > out_len) FAIL_BUF;
                            wr16(out + op, len); op += 2;
                        } else {
                            if (op + 6 > out_len) FAIL_BUF;
                            wr16(out + op, 0); wr32(out + op + 2, len); op += 6;
                        }
                    }
                }
            }
            flags |= 1;
        }
        if (++flag_count == 32) {                                                 /* :317-324 */
            wr32(out + flags_at, flags);
            flag_count = 0;
            if (op + 4 > out_len) FAIL_BUF;
            flags_at = op; op += 4;
        }
    }
    while (ip < in_len && op < out_len) {                                         /* :326-338 tail literals */
        out[op++] = in[ip++];
        flags <<= 1;
        if (++flag_count == 32) {
            wr32(out + flags_at, flags);
            flag_count = 0;
            if (op + 4 > out_len) FAIL_BUF;
            flags_at = op; op += 4;
        }
    }
    if (ip != in_len) FAIL_BUF;
    flags = flag_count ? (flags << (32 - flag_count)) | ((1u << (32 - flag_count)) - 1) : 0xFFFFFFFFu;   /* :343 */
    wr32(out + flags_at, flags);
    xd_free(&d);
    *out_len_p = op;
    return ORC_OK;
#undef FAIL_BUF
}
