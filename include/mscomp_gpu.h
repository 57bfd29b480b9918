/*
 * mscomp_gpu.h -- device-resident batch interface of the B200-native libMSCompression (not in the reference).
 *
 * The reference is host-memory-in / host-memory-out (include/mscomp.h:59,:88); the six one-shot functions of this
 * library (lznt1.h, xpress.h, xpress_huff.h) are thin wrappers that stage host buffers through the functions
 * below.  Callers that already hold their data in HBM use these directly: all `d_` pointers are DEVICE pointers,
 * every call is asynchronous on the context's CUDA stream, results (sizes / statuses) are written to device
 * memory and are valid after mscomp_gpu_synchronize() or any later work on the same stream.
 *
 * Data-parallel units: LZNT1 -- the 4 KiB chunk; Xpress-Huffman and plain Xpress -- the
 * independent buffer (its own stream, as WIM stores them).  Nothing here falls back to the CPU.
 *
 * Alignment: kernels read 16-byte granules that contain at least one valid byte, so buffers may start and end at
 * any address; best throughput with 16-byte aligned bases.
 */
#ifndef MSCOMP_GPU_H
#define MSCOMP_GPU_H
#include "mscomp/general.h"

EXTERN_C_START

typedef struct mscomp_gpu_ctx mscomp_gpu_ctx;

/* Summary written by the stream-level calls. */
typedef struct mscomp_gpu_result {
    uint64_t bytes;        /* total bytes produced */
    int32_t  status;       /* MSCOMP_OK or the status of the first failing unit */
    uint32_t first_bad;    /* index of the first failing unit (valid when status < 0) */
    uint32_t short_units;  /* LZNT1 decode: number of non-final chunks that produced fewer than 4096 bytes */
    uint32_t reserved;
} mscomp_gpu_result;

/* Per-buffer descriptors of a batch (device arrays of length count). */
typedef struct mscomp_gpu_batch {
    uint32_t        count;
    const uint64_t* d_in_off;    /* byte offset of buffer i inside d_in */
    const uint32_t* d_in_len;    /* its length */
    const uint64_t* d_out_off;   /* where its output starts inside d_out */
    const uint32_t* d_out_cap;   /* output capacity */
    uint32_t*       d_out_len;   /* OUT: bytes produced */
    int32_t*        d_status;    /* OUT: MSCompStatus per buffer */
} mscomp_gpu_batch;

/* mscomp_gpu_create: the context creates and owns a non-blocking stream.
 * mscomp_gpu_create_on_stream: all work is issued on the caller's stream (a cudaStream_t passed as void*; NULL is
 * the legacy default stream), so it orders with the caller's other work on that stream.
 * Both return NULL when the device cannot be used. */
MSCOMPAPI mscomp_gpu_ctx* mscomp_gpu_create(int device);
MSCOMPAPI mscomp_gpu_ctx* mscomp_gpu_create_on_stream(int device, void* cuda_stream);
MSCOMPAPI void            mscomp_gpu_destroy(mscomp_gpu_ctx* ctx);
MSCOMPAPI MSCompStatus    mscomp_gpu_synchronize(mscomp_gpu_ctx* ctx);
/* Kernels launched through this context so far (the benchmark reports it as gpu_launches). */
MSCOMPAPI uint64_t        mscomp_gpu_kernel_launches(const mscomp_gpu_ctx* ctx);
/* Optional per-kernel timing: CUDA events on the context's stream around every launch (at most 2048 launches between
 * reports).  mscomp_gpu_profile_report synchronises and writes one "kernel_name calls total_ms" line per kernel. */
MSCOMPAPI void            mscomp_gpu_profile_enable(mscomp_gpu_ctx* ctx, int on);
MSCOMPAPI size_t          mscomp_gpu_profile_report(mscomp_gpu_ctx* ctx, char* buf, size_t cap);
/* Xpress / Xpress-Huffman encoders of this context.  MSCOMP_GPU_ENCODER_REFERENCE (the default): the reference's own parse --
 * hash chains of 11 links, nice length 48 (include/mscomp/XpressDictionary.h:28-36, level 3) and, for Xpress-Huffman, its
 * heap-built code lengths: the streams are byte-identical to xpress_compress / xpress_huff_compress of the reference.
 * MSCOMP_GPU_ENCODER_FAST: a bulk-synchronous matcher with two candidates per position (and a window reset per 64 KiB chunk for
 * Xpress-Huffman): valid streams for the reference's decoders, 2-3x the encode speed, 5-9 % larger on text.
 * The initial mode of a context is REFERENCE unless the environment has MSCOMP_GPU_ENCODER=fast.  LZNT1 has one encoder. */
#define MSCOMP_GPU_ENCODER_REFERENCE 0
#define MSCOMP_GPU_ENCODER_FAST      1
MSCOMPAPI void            mscomp_gpu_set_encoder(mscomp_gpu_ctx* ctx, int mode);
MSCOMPAPI int             mscomp_gpu_get_encoder(const mscomp_gpu_ctx* ctx);

/* ---- LZNT1: one logical stream of 4 KiB chunks ------------------------------------------------------------ */
/* Device path behind lznt1_compress (reference src/lznt1_compress.cpp:233-273).  out_cap must be at least
 * lznt1_max_compressed_size(in_len).  d_chunk_off (may be NULL) receives ceil(in_len/4096)+1 offsets: the byte
 * offset of every chunk header in d_out, then the total.  d_result->bytes = total compressed bytes. */
MSCOMPAPI MSCompStatus mscomp_gpu_lznt1_compress(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len,
                                                 uint8_t* d_out, size_t out_cap, uint64_t* d_chunk_off,
                                                 mscomp_gpu_result* d_result);
/* Device path behind lznt1_decompress (reference src/lznt1_decompress.cpp:122-209, :37-121).  d_chunk_off[i] is
 * the offset of chunk i's 2-byte header in d_in (the format stores no index: take it from
 * mscomp_gpu_lznt1_compress, or from a walk over the headers as the one-shot wrapper does).  Chunk i is written to
 * d_out + 4096*i, so d_out needs 4096*nchunks bytes; the stream is contiguous when d_result->short_units == 0
 * (always true for streams made by the one-shot compressors).  d_chunk_size / d_chunk_status (may be NULL)
 * receive the per-chunk byte counts and statuses. */
MSCOMPAPI MSCompStatus mscomp_gpu_lznt1_decompress(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len,
                                                   const uint64_t* d_chunk_off, uint32_t nchunks, uint8_t* d_out,
                                                   uint32_t* d_chunk_size, int32_t* d_chunk_status,
                                                   mscomp_gpu_result* d_result);

/* Chunk index of an unindexed LZNT1 stream, built on the GPU (the format stores none; the reference walks the headers
 * serially: src/lznt1_decompress.cpp:254-273 with the header read of :127-135).  Synchronous.  d_chunk_off receives the
 * offset of every chunk header in stream order (chunk_cap entries available).  h_result (host memory): number of chunks,
 * bytes consumed by whole chunks, and how the chain ended: 0 clean end, 1 one stray zero byte, 2 one stray non-zero
 * byte, 3 two-byte terminator at the very end, 4 terminator followed by data, 5 truncated chunk, 6 bad signature -- the
 * conditions the reference's driver turns into MSCOMP_OK / BUF_ERROR / DATA_ERROR.  Returns MSCOMP_BUF_ERROR (with
 * h_result filled) when the stream has more than chunk_cap chunks. */
typedef struct mscomp_gpu_lznt1_index_result { uint64_t nchunks; uint64_t consumed; int32_t tail; int32_t reserved; } mscomp_gpu_lznt1_index_result;
MSCOMPAPI MSCompStatus mscomp_gpu_lznt1_index(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len,
                                              uint64_t* d_chunk_off, size_t chunk_cap,
                                              mscomp_gpu_lznt1_index_result* h_result);

/* ---- Xpress-Huffman and plain Xpress: batches of independent buffers ----------------------------------------- */
/* Device paths behind xpress_huff_compress / _decompress (reference src/xpress_huff_compress.cpp:247-331,
 * src/xpress_huff_decompress.cpp:136-162) and xpress_compress / _decompress (src/xpress_compress.cpp:240-347,
 * src/xpress_decompress.cpp:405-462).  Each buffer is a complete stream of the format. */
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_huff_compress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                             const mscomp_gpu_batch* batch);
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_huff_decompress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                               const mscomp_gpu_batch* batch);
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_compress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                        const mscomp_gpu_batch* batch);
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_decompress_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                          const mscomp_gpu_batch* batch);

/* An Xpress-Huffman stream is bit-serial (variable-length codes, raw length bytes at the reader's byte cursor, chunk ends stored
 * nowhere: reference src/xpress_huff_decompress.cpp:39-162), so without further knowledge one GPU lane walks a stream.  A
 * side-band RESTART INDEX -- never part of the stream -- records the reference reader's state (include/mscomp/Bitstream.h:34-44)
 * at every 64th symbol; with it every 64-symbol segment decodes on its own lane (one warp per stream).
 *   mscomp_gpu_xpress_huff_decompress_batch          builds the index in scratch memory (serial walk, one lane per stream), then decodes
 *   mscomp_gpu_xpress_huff_index_batch               only the walk: fills out_len / status of the batch (no output bytes) and the index
 *   mscomp_gpu_xpress_huff_decompress_batch_indexed  only the decode, from an index kept by the caller (e.g. made once after compressing)
 * Stream i owns entries d_index[d_index_off[i] .. d_index_off[i+1]) (count+1 offsets, in entries); d_nseg[i] = entries in use.
 * mscomp_gpu_xpress_huff_index_bound(out_cap) entries are always enough for a stream with out_cap bytes of output room.  The
 * decoder follows an index only as far as it is consistent with the stream: every segment must end exactly in the next entry's
 * state; streams where it does not are decoded without the index (same bytes, same status codes as the reference). */
typedef struct mscomp_gpu_xhuff_seg {
    uint32_t mask;        /* bit window of the reader, valid bits left-aligned */
    uint32_t ip;          /* byte offset of the next unread byte of the stream */
    uint32_t out_pos;     /* output bytes produced before the segment's first symbol */
    uint32_t info;        /* bits 0-5: valid bits in mask; bit 7: first segment of a chunk (its 256-byte table is at ip - 260);
                             bits 8-15: symbols in the segment (1..64) */
} mscomp_gpu_xhuff_seg;
MSCOMPAPI uint64_t     mscomp_gpu_xpress_huff_index_bound(uint64_t out_cap);
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_huff_index_batch(mscomp_gpu_ctx* ctx, const uint8_t* d_in, const mscomp_gpu_batch* batch,
                                                          mscomp_gpu_xhuff_seg* d_index, const uint64_t* d_index_off, uint32_t* d_nseg);
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_huff_decompress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                                       const mscomp_gpu_batch* batch, const mscomp_gpu_xhuff_seg* d_index,
                                                                       const uint64_t* d_index_off, const uint32_t* d_nseg);
/* Compress AND leave the restart index of every stream (room: mscomp_gpu_xpress_huff_index_bound(in_len[i]) entries for stream i):
 * the streams are byte for byte those of mscomp_gpu_xpress_huff_compress_batch; the index is what the walk would find. */
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_huff_compress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                                     const mscomp_gpu_batch* batch, mscomp_gpu_xhuff_seg* d_index,
                                                                     const uint64_t* d_index_off, uint32_t* d_nseg);
/* Diagnostic: streams the indexed decoder handed to the index-free decoder since the context was created (synchronises). */
MSCOMPAPI uint64_t     mscomp_gpu_xpress_huff_second_chances(mscomp_gpu_ctx* ctx);

/* Plain Xpress is one unbroken stream per buffer (flag words and shared length nibbles chain every item to the previous
 * ones), so a reference-made stream decodes on one GPU lane.  Streams made HERE can carry a side-band restart index --
 * never part of the stream, the stream stays byte-for-byte what the call above produces -- with one entry per seg_bytes
 * (a multiple of 2048) of input: the decoder state at an item boundary.  With it the decoder parses every segment on its own
 * lane.  d_index holds index_stride entries per buffer (entry 0 carries the number of entries actually used; buffers with
 * more segments than index_stride get longer last segments). */
typedef struct mscomp_gpu_xpress_seg {
    uint32_t in_pos;      /* compressed offset of the next byte to read */
    uint32_t out_pos;     /* uncompressed position of the next item */
    uint32_t flag_pos;    /* compressed offset of the flag word in force (when flag_used != 0) */
    uint32_t flag_used;   /* items of that flag word already consumed, 0..31 */
    uint32_t half_pos;    /* compressed offset of the byte whose high length nibble is still unused, 0xFFFFFFFF = none */
    uint32_t nseg;        /* entry 0 only: entries of this buffer */
} mscomp_gpu_xpress_seg;
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_compress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                                const mscomp_gpu_batch* batch, uint32_t seg_bytes,
                                                                mscomp_gpu_xpress_seg* d_index, uint32_t index_stride);
MSCOMPAPI MSCompStatus mscomp_gpu_xpress_decompress_batch_indexed(mscomp_gpu_ctx* ctx, const uint8_t* d_in, uint8_t* d_out,
                                                                  const mscomp_gpu_batch* batch,
                                                                  const mscomp_gpu_xpress_seg* d_index, uint32_t index_stride);

/* ---- several GPUs: one process (or thread) and one context per GPU ------------------------------------------------------- */
/* The units of all three formats are independent, so a job over N GPUs is N single-GPU jobs on contiguous unit ranges plus ONE
 * exchange -- every rank learns every rank's output sizes (an ncclAllGather on the context's stream, nothing else crosses
 * NVLink).  NCCL is bound at run time (libnccl.so.2); single-GPU users need none.
 *   mscomp_gpu_shard_range      the unit range of a rank: [floor(rank * units / nranks), floor((rank + 1) * units / nranks))
 *   mscomp_gpu_comm_unique_id   rank 0 makes the 128-byte id and hands it to the other ranks by its own means (MPI, a file, a socket)
 *   mscomp_gpu_comm_init        collective: every rank, with that id (ncclCommInitRank); the communicator belongs to the context
 *   mscomp_gpu_comm_attach      instead: use an ncclComm_t the caller made (passed as void*); it is not destroyed with the context
 *   mscomp_gpu_sizes_allgather  d_all[r * n_values + k] = rank r's d_mine[k] (device uint64 arrays; d_all holds nranks * n_values)
 *   mscomp_gpu_lznt1_compress_sharded / mscomp_gpu_compress_batch_sharded: this rank's shard through the single-GPU path, then
 *       d_sizes_all[r] = compressed bytes of rank r -- the global stream / batch is the ranks' parts in rank order. */
#define MSCOMP_GPU_COMM_ID_BYTES 128
MSCOMPAPI void         mscomp_gpu_shard_range(uint64_t units, uint32_t rank, uint32_t nranks, uint64_t* first, uint64_t* count);
MSCOMPAPI MSCompStatus mscomp_gpu_comm_unique_id(void* id128);
MSCOMPAPI MSCompStatus mscomp_gpu_comm_init(mscomp_gpu_ctx* ctx, const void* id128, uint32_t rank, uint32_t nranks);
MSCOMPAPI MSCompStatus mscomp_gpu_comm_attach(mscomp_gpu_ctx* ctx, void* nccl_comm, uint32_t rank, uint32_t nranks);
MSCOMPAPI void         mscomp_gpu_comm_release(mscomp_gpu_ctx* ctx);
MSCOMPAPI uint32_t     mscomp_gpu_comm_rank(const mscomp_gpu_ctx* ctx);
MSCOMPAPI uint32_t     mscomp_gpu_comm_size(const mscomp_gpu_ctx* ctx);
MSCOMPAPI MSCompStatus mscomp_gpu_sizes_allgather(mscomp_gpu_ctx* ctx, const uint64_t* d_mine, uint32_t n_values, uint64_t* d_all);
MSCOMPAPI uint64_t     mscomp_gpu_collectives(const mscomp_gpu_ctx* ctx);     /* NCCL collectives issued through this context */
MSCOMPAPI MSCompStatus mscomp_gpu_lznt1_compress_sharded(mscomp_gpu_ctx* ctx, const uint8_t* d_in, size_t in_len, uint8_t* d_out, size_t out_cap,
                                                         uint64_t* d_chunk_off, mscomp_gpu_result* d_result, uint64_t* d_sizes_all);
MSCOMPAPI MSCompStatus mscomp_gpu_compress_batch_sharded(mscomp_gpu_ctx* ctx, int format, const uint8_t* d_in, uint8_t* d_out,
                                                         const mscomp_gpu_batch* batch, uint64_t* d_my_total, uint64_t* d_sizes_all);

EXTERN_C_END
#endif
