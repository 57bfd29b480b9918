// xhuff_index.cuh -- restart index of Xpress-Huffman streams: the reader state at every 64th symbol.
//
// An Xpress-Huffman stream is bit-serial: variable-length codes, raw length bytes interleaved at the reader's byte
// cursor, chunk boundaries stored nowhere (reference src/xpress_huff_decompress.cpp:39-162, include/mscomp/Bitstream.h:34-107;
// SURVEY.md 0.2).  What CAN be recorded is the exact state of the reference's reader -- bit window, valid bits, byte
// cursor, output position -- at chosen symbol boundaries.  From such a state a decoder can start in the middle of the
// stream, so the segments between two recorded states decode in parallel (xhuff_decode_indexed.cuh: one LANE per
// segment).  The index is side-band: it is never part of the stream.
//
// xhuff_index_kernel builds it for streams that come without one: one lane per stream walks the stream with the
// reference's reader (xh_step of xhuff_decode.cuh), writes no output bytes, and emits one entry per 64 symbols plus the
// stream's status and output length -- it is also the authority on the status code of a damaged stream.
#pragma once
#include "xhuff_decode.cuh"

namespace msc {

constexpr uint32_t XHI_SEG_SYMS = 64;          // symbols per segment (the last segment of a chunk may hold fewer)
constexpr uint32_t XHI_NEWCHUNK = 0x80u;       // info flag: this segment is the first of a chunk; its code-length table sits at ip - 260

// One index entry = mscomp_gpu_xhuff_seg (include/mscomp_gpu.h).
struct XhSeg {
    uint32_t mask;       // bit window of the reader (InputBitstream::mask), valid bits left-aligned
    uint32_t ip;         // byte cursor: offset of the next unread byte of the stream
    uint32_t out_pos;    // output bytes produced before the segment's first symbol
    uint32_t info;       // bits 0-5: valid bits in mask (0..32); bit 7: XHI_NEWCHUNK; bits 8-15: symbols in the segment (1..64)
};

// Entries a stream with `out_cap` bytes of output room can need: every symbol yields at least one byte, every chunk but
// the last at least 65,536, and every chunk ends with one possibly short segment.
__host__ __device__ inline uint64_t xhi_max_segments(uint64_t out_cap) { return out_cap / XHI_SEG_SYMS + out_cap / XH_CHUNK + 3; }

// Sink of the lane parser that keeps reader states instead of bytes.
struct LaneIdx {
    XhSeg* ent;          // this stream's entries
    uint32_t cap, nent;  // room / entries needed so far
    uint32_t opos;       // output bytes the symbols so far stand for
    uint32_t cur;        // symbols counted in the open segment
    XhSeg open;          // the open segment's entry
    bool is_open, fresh; // a segment is open / the next segment to open is the first of its chunk
    __device__ __forceinline__ void init(XhSeg* e, uint32_t c)
    { ent = e; cap = c; nent = 0; opos = 0; cur = 0; is_open = false; fresh = false; open = XhSeg{0, 0, 0, 0}; }
    __device__ __forceinline__ void close()
    {
        if (is_open && cur) {                          // (a segment that saw no symbol -- only the end-of-stream code -- leaves no entry)
            open.info |= cur << 8;
            if (nent < cap) ent[nent] = open;
            ++nent;
        }
        is_open = false; cur = 0;
    }
    __device__ __forceinline__ void new_chunk() { close(); fresh = true; }
    // called with the reader state right before every symbol is decoded
    __device__ __forceinline__ void before_symbol(uint32_t mask, uint32_t bits, uint32_t ip)
    {
        if (!is_open) { open = XhSeg{mask, ip, opos, bits | (fresh ? XHI_NEWCHUNK : 0u)}; is_open = true; fresh = false; }
    }
    // called for every symbol that produces output (not for the end-of-stream code)
    __device__ __forceinline__ void emit(bool, uint32_t, uint32_t, uint32_t len)
    {
        opos += len;
        if (++cur == XHI_SEG_SYMS) close();
    }
    __device__ __forceinline__ void finish() { close(); }
};

// Stream i = in[in_off[i] .. +in_len[i]); its entries go to index + index_off[i] (room: index_off[i+1] - index_off[i]).
// out_len[i] / status[i] are what the one-shot decoder would return for capacity out_cap[i]; nseg[i] = entries needed
// (a stream whose entries did not fit gets status MSCOMP_MEM_ERROR: the caller sized the index too small).
__global__ void __launch_bounds__(32 * XHD_WARPS_PER_BLOCK)
xhuff_index_kernel(const uint8_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
                   const uint32_t* __restrict__ out_cap, XhSeg* __restrict__ index, const uint64_t* __restrict__ index_off,
                   uint32_t* __restrict__ nseg, uint32_t* __restrict__ out_len, int32_t* __restrict__ status)
{
    MSC_DYN_SMEM(smem_raw);
    XhWarpSmem& sm = reinterpret_cast<XhWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const int lane = lane_id();
    const uint32_t warps_total = gridDim.x * XHD_WARPS_PER_BLOCK;
    for (uint32_t g = blockIdx.x * XHD_WARPS_PER_BLOCK + (threadIdx.x >> 5); g * 32 < count; g += warps_total) {
        const uint32_t i = g * 32 + lane;
        XhLane L;
        L.status = 1; L.in = in; L.in_len = 0; L.cap = 0; L.ip = 0; L.nused = 0;
        L.mask = 0; L.bits = 0; L.chunk_end = 0; L.need_table = true; L.spec = false; L.reach = 0;
#pragma unroll
        for (int k = 0; k < 15; ++k) L.lims[k] = 0x8000u;
#pragma unroll
        for (int k = 0; k < 8; ++k) { L.baseP[k] = 0; L.litendP[k] = 0; }
        LaneIdx O;
        O.init(index, 0);
        if (i < count) {
            L.in = in + in_off[i]; L.in_len = in_len[i]; L.cap = out_cap[i];
            const uint64_t room = index_off[i + 1] - index_off[i];
            O.init(index + index_off[i], room > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)room);
            L.status = 0;
        }
        L.win.init(L.in, L.in + L.in_len);
        while (__any_sync(FULL, L.status == 0)) {
            if (L.status == 0) xh_step(L, sm, lane, O);
        }
        O.finish();
        if (i < count) {
            int st = L.status == 1 ? ST_OK : L.status;
            if (st == ST_OK && O.nent > O.cap) st = ST_MEM;
            out_len[i] = O.opos; status[i] = st; nseg[i] = O.nent;
        }
    }
}

} // namespace msc
