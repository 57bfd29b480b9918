/*
 * mscomp/general.h -- types shared by every header of the B200-native libMSCompression.
 *
 * Drop-in for the reference's include/mscomp/general.h: same type names, same enumerator values
 * (formats general.h:65-71, statuses :74-85, flush modes :88-93 of coderforlife/ms-compress), so code written
 * against the reference recompiles unchanged.  The text is new; nothing is copied from the reference (GPLv3).
 */
#ifndef MSCOMP_GENERAL_H
#define MSCOMP_GENERAL_H

#include <stddef.h>
#include <stdint.h>

#if defined(_WIN32) && defined(MSCOMP_API_DLL)
#  ifdef MSCOMP_API_EXPORT
#    define MSCOMPAPI __declspec(dllexport)
#  else
#    define MSCOMPAPI __declspec(dllimport)
#  endif
#elif defined(__GNUC__)
#  define MSCOMPAPI __attribute__((visibility("default")))
#else
#  define MSCOMPAPI
#endif

#ifdef __cplusplus
#  define EXTERN_C_START extern "C" {
#  define EXTERN_C_END   }
#else
#  define EXTERN_C_START
#  define EXTERN_C_END
#endif

/* byte-pointer aliases used throughout the reference's prototypes */
typedef uint8_t        byte;
typedef byte*          bytes;
typedef const byte     const_byte;
typedef const_byte*    const_bytes;

/* Compression formats; the numbers are the Windows COMPRESSION_FORMAT_* constants. */
typedef enum _MSCompFormat {
    MSCOMP_NONE        = 0,
    MSCOMP_RESERVED    = 1,   /* never valid */
    MSCOMP_LZNT1       = 2,
    MSCOMP_XPRESS      = 3,
    MSCOMP_XPRESS_HUFF = 4
} MSCompFormat;

/* Result codes: zero = success, positive = stream information, negative = failure. */
typedef enum _MSCompStatus {
    MSCOMP_OK                  = 0,
    MSCOMP_STREAM_END          = 1,
    MSCOMP_POSSIBLE_STREAM_END = 2,
    MSCOMP_ERRNO               = -1,
    MSCOMP_ARG_ERROR           = -2,   /* bad format / bad argument */
    MSCOMP_DATA_ERROR          = -3,   /* malformed compressed data */
    MSCOMP_MEM_ERROR           = -4,   /* host or device allocation failed, or no usable CUDA device */
    MSCOMP_BUF_ERROR           = -5    /* output buffer too small / input truncated */
} MSCompStatus;

typedef enum _MSCompFlush {
    MSCOMP_NO_FLUSH = 0,
    MSCOMP_FLUSH    = 2,
    MSCOMP_FINISH   = 4
} MSCompFlush;

/* Stream object of the *_deflate / *_inflate calls; same layout as the reference's (include/mscomp/general.h:96-120, built with
 * its default configuration: error and warning texts present).  The caller sets in / in_avail / out / out_avail between calls;
 * the library advances them and keeps the totals. */
typedef struct _mscomp_internal_state mscomp_internal_state;
typedef struct _mscomp_stream {
    MSCompFormat format;
#ifdef __cplusplus
    bool         compressing;
#else
    int          compressing;
#endif
    const_bytes  in;          /* next input byte */
    size_t       in_avail;    /* bytes available at in */
    size_t       in_total;    /* input bytes consumed so far */
    bytes        out;         /* next output byte goes here */
    size_t       out_avail;   /* free space at out */
    size_t       out_total;   /* bytes produced so far */
    char         error[256];
    char         warning[256];
    mscomp_internal_state* state;
} mscomp_stream;

#endif
