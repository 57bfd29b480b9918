// emu_kernels.cpp -- TEST INFRASTRUCTURE: compiles the CUDA kernel sources for the CPU through simt_emu.h and
// exposes one C entry point per kernel so that pytest can drive them with host buffers (see tests/test_emu_*.py).
#include "simt_emu.h"
#include "../../ms-compress_b200/csrc/kernels/lznt1_decode.cuh"
#include "../../ms-compress_b200/csrc/kernels/lznt1_encode.cuh"
#include "../../ms-compress_b200/csrc/kernels/lznt1_index.cuh"
#include "../../ms-compress_b200/csrc/kernels/compact.cuh"
#include "../../ms-compress_b200/csrc/kernels/xhuff_decode.cuh"
#include "../../ms-compress_b200/csrc/kernels/xhuff_index.cuh"
#include "../../ms-compress_b200/csrc/kernels/xhuff_decode_indexed.cuh"
#include "../../ms-compress_b200/csrc/kernels/xhuff_chunks.cuh"
#include "../../ms-compress_b200/csrc/kernels/xhuff_chunk_warp.cuh"
#include "../../ms-compress_b200/csrc/kernels/lz_jump.cuh"
#include "../../ms-compress_b200/csrc/kernels/xpress_decode.cuh"
#include "../../ms-compress_b200/csrc/kernels/xhuff_encode.cuh"
#include "../../ms-compress_b200/csrc/kernels/xpress_encode.cuh"
#include "../../ms-compress_b200/csrc/kernels/xpress_spec.cuh"
#include <vector>
#include <algorithm>

using emu::dim3e;

extern "C" {

void emu_lznt1_decode(const uint8_t* in, uint64_t in_len, const uint64_t* chunk_off, uint32_t nchunks,
                      uint8_t* out, uint32_t* out_size, int32_t* status, uint32_t grid)
{
    emu::launch(msc::lznt1_decode_kernel, dim3e(grid), dim3e(msc::LZNT1_WARPS_PER_BLOCK * 32),
                in, in_len, chunk_off, nchunks, out, out_size, status);
}

// whole-stream LZNT1 compress: encode kernel -> size scan -> gather.  `slots` = nchunks*4112 scratch.
void emu_lznt1_encode(const uint8_t* in, uint64_t in_len, uint32_t nchunks, uint8_t* slots, uint32_t* slot_size,
                      uint64_t* offs, uint8_t* out, uint32_t grid, uint32_t heavy_work)
{
    emu::launch(msc::lznt1_encode_kernel, dim3e(grid), dim3e(msc::LZE_THREADS), in, in_len, 0u, nchunks, slots, slot_size, heavy_work);
    emu::launch(msc::scan_sizes_kernel, dim3e(1), dim3e(msc::SCAN_THREADS), (const uint32_t*)slot_size, nchunks, (uint64_t)0, (const uint64_t*)nullptr, offs, (uint64_t*)nullptr);
    emu::launch(msc::gather_slots_kernel, dim3e(grid), dim3e(256), (const uint8_t*)slots, (uint64_t)msc::LZNT1_SLOT,
                (const uint32_t*)slot_size, (const uint64_t*)offs, nchunks, out);
}


void emu_xhuff_decode(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                      const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    emu::launch(msc::xhuff_decode_kernel, dim3e(grid), dim3e(32 * msc::XHD_WARPS_PER_BLOCK), in, out, count, in_off, in_len, out_off, out_cap, out_len, status);
}


void emu_xpress_decode(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                       const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    emu::launch(msc::xpress_decode_kernel, dim3e(grid), dim3e(32 * msc::XPD_WARPS_PER_BLOCK), in, out, count, in_off, in_len, out_off, out_cap, out_len, status);
}


void emu_xhuff_encode(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                      const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    std::vector<uint8_t> scratch((size_t)grid * msc::XHE_SCRATCH_BYTES + 64);
    emu::launch(msc::xhuff_encode_kernel, dim3e(grid), dim3e(msc::XHE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                scratch.data(), 0, (msc::XhSeg*)nullptr, (const uint64_t*)nullptr, (uint32_t*)nullptr);
}
// the reference's LZ77 tokens: links, Find, assumed parse per tile, then the encoder.  eos_mode 1: the buffers are the consecutive 64 KiB
// chunks of ONE stream (Find sees the stream); eos_mode 0: independent buffers (of any length).  With `index` the encoder also writes
// its restart index (eos_mode 0), laid out like emu_xhuff_encode_index_check's.
static void emu_xhuff_exact_front(const uint8_t* in, uint32_t count, const uint64_t* in_off, const uint32_t* in_len, uint32_t grid, int eos_mode,
                                  std::vector<uint32_t>& seg_first, std::vector<uint32_t>& lenoff, std::vector<uint32_t>& visited0,
                                  std::vector<uint32_t>& exit0, std::vector<uint32_t>& desc_len, std::vector<uint16_t>& dist_out)
{
    using namespace msc;
    unsigned long long totals[3] = {0, 0, 0};
    uint32_t work[4] = {0, 0, 0, 0};
    std::vector<uint64_t> desc_off(1);
    seg_first.assign(count + 2, 0);
    desc_len.assign(1, 0);
    uint32_t fcount = count;
    const uint64_t* f_off = in_off;
    const uint32_t* f_len = in_len;
    bool big;
    if (eos_mode == 1) {
        emu::launch(xhuff_stream_desc_kernel, dim3e(1), dim3e(32), count, in_off, in_len, desc_off.data(), desc_len.data(), seg_first.data(), work);
        totals[0] = count; big = count > 1; fcount = 1; f_off = desc_off.data(); f_len = desc_len.data();
    } else {
        emu::launch(xpress_segments_kernel, dim3e(1), dim3e(1024), in_len, count, seg_first.data(), totals, work);
        big = totals[2] > XHE_CHUNK;
    }
    const size_t npos = (size_t)totals[0] * LZX_SEG;
    lenoff.assign(npos + 1, LZX_UNKNOWN); visited0.assign(npos / 32 + 1, 0xDEADBEEFu); exit0.assign(npos / LZF_TILE + 1, 0xDEADBEEFu);
    std::vector<uint32_t> exit_sub(npos / LZX_SUB + 1, 0xDEADBEEFu);
    dist_out.assign(npos + 1, 0xDEAD);
    if (!totals[0]) return;
    if (big) emu::launch(lz_prev_kernel<uint32_t>, dim3e(grid), dim3e(32), in, fcount, f_off, f_len, (const uint32_t*)seg_first.data(), work, dist_out.data(), 65535u, LZX_SEG);
    else emu::launch(lz_prev_kernel<uint16_t>, dim3e(grid), dim3e(32), in, fcount, f_off, f_len, (const uint32_t*)seg_first.data(), work, dist_out.data(), 65535u, LZX_SEG);
    emu::launch(lz_find_kernel<LZX_FIND_MINB>, dim3e(grid), dim3e(LZX_FIND_THREADS), in, fcount, f_off, f_len, (const uint32_t*)seg_first.data(), work + 1,
                (const uint16_t*)dist_out.data(), lenoff.data(), visited0.data(), exit_sub.data(), 65535u, XHE_CHUNK);
    emu::launch(lz_stitch_kernel, dim3e(grid), dim3e(128), in, fcount, f_off, f_len, (const uint32_t*)seg_first.data(), (const uint16_t*)dist_out.data(),
                lenoff.data(), visited0.data(), (const uint32_t*)exit_sub.data(), exit0.data(), 65535u, XHE_CHUNK);
}
void emu_xhuff_encode_exact(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                            const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid, uint32_t eos_mode)
{
    using namespace msc;
    std::vector<uint32_t> seg_first, lenoff, visited0, exit0, desc_len;
    std::vector<uint16_t> dist;
    emu_xhuff_exact_front(in, count, in_off, in_len, grid, (int)eos_mode, seg_first, lenoff, visited0, exit0, desc_len, dist);
    std::vector<uint8_t> scratch((size_t)grid * XHE_SCRATCH_BYTES + 64);
    XhPre pre{seg_first.data(), lenoff.data(), visited0.data(), exit0.data(), dist.data(), desc_len.data()};
    emu::launch(xhuff_encode_exact_kernel, dim3e(grid), dim3e(XHE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                scratch.data(), (int)eos_mode, (XhSeg*)nullptr, (const uint64_t*)nullptr, (uint32_t*)nullptr, pre);
}
// the encoder's own restart index next to the one the serial walk finds for the streams it wrote: returns the number of streams
// whose two indices differ (entry count or any entry)
uint32_t emu_xhuff_encode_index_check(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                                      const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    using namespace msc;
    std::vector<uint8_t> scratch((size_t)grid * XHE_SCRATCH_BYTES + 64);
    std::vector<uint64_t> ioff(count + 1, 0);
    for (uint32_t i = 0; i < count; ++i) ioff[i + 1] = ioff[i] + xhi_max_segments(in_len[i]);
    std::vector<XhSeg> mine(ioff[count] + 1), walk(ioff[count] + 1);
    std::vector<uint32_t> n_mine(count + 1), n_walk(count + 1), wl(count + 1);
    std::vector<int32_t> ws(count + 1);
    emu::launch(xhuff_encode_kernel, dim3e(grid), dim3e(XHE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                scratch.data(), 0, mine.data(), (const uint64_t*)ioff.data(), n_mine.data());
    emu::launch(xhuff_index_kernel, dim3e(grid), dim3e(32 * XHD_WARPS_PER_BLOCK), (const uint8_t*)out, count, out_off, (const uint32_t*)out_len, in_len,
                walk.data(), (const uint64_t*)ioff.data(), n_walk.data(), wl.data(), ws.data());
    uint32_t bad = 0;
    for (uint32_t i = 0; i < count; ++i) {
        bool same = status[i] == 0 && ws[i] == 0 && wl[i] == in_len[i] && n_mine[i] == n_walk[i];
        for (uint32_t e = 0; same && e < n_mine[i]; ++e) {
            const XhSeg &a = mine[ioff[i] + e], &b = walk[ioff[i] + e];
            same = a.mask == b.mask && a.ip == b.ip && a.out_pos == b.out_pos && a.info == b.info;
        }
        bad += same ? 0 : 1;
    }
    return bad;
}


void emu_xpress_encode(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                       const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    emu::launch(msc::xpress_encode_kernel, dim3e(grid), dim3e(msc::XPE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                0u, (msc::XpressSeg*)nullptr, 0u);
}


// two-kernel decoders: parse (one lane per buffer, tokens) + apply (one warp per buffer)
static uint64_t emu_extent(uint32_t count, const uint64_t* out_off, const uint32_t* out_cap)
{ uint64_t e = 0; for (uint32_t i = 0; i < count; ++i) e = std::max<uint64_t>(e, out_off[i] + out_cap[i]); return e; }

void emu_xhuff_decode2(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                       const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    std::vector<uint32_t> tok(emu_extent(count, out_off, out_cap) + 16), ntok(count + 1);
    emu::launch(msc::xhuff_parse_kernel, dim3e(grid), dim3e(32 * msc::XHD_WARPS_PER_BLOCK), in, tok.data(), count, in_off, in_len, out_off, out_cap, out_len, status, ntok.data());
    emu::launch(msc::lz_apply_kernel, dim3e(grid), dim3e(32 * msc::LZA_WARPS_PER_BLOCK), (const uint32_t*)tok.data(), (const uint32_t*)ntok.data(), out_off, (const int32_t*)status, count, out);
}
void emu_xpress_decode2(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                        const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    std::vector<uint32_t> tok(emu_extent(count, out_off, out_cap) + 16), ntok(count + 1);
    emu::launch(msc::xpress_parse_kernel, dim3e(grid), dim3e(32 * msc::XPD_WARPS_PER_BLOCK), in, tok.data(), count, in_off, in_len, out_off, out_cap, out_len, status, ntok.data());
    emu::launch(msc::lz_apply_kernel, dim3e(grid), dim3e(32 * msc::LZA_WARPS_PER_BLOCK), (const uint32_t*)tok.data(), (const uint32_t*)ntok.data(), out_off, (const int32_t*)status, count, out);
}

// Xpress-Huffman through the restart index: serial walk (one lane per stream) -> index; decode (one lane per segment); then the
// second-chance pass for streams the indexed decoder gave up on.  damage != 0: entry `damage-1` of every index gets its byte
// cursor shifted before decoding (an index that does not belong to the stream).  Returns the number of second chances taken.
static uint64_t g_xhuff_second = 0;
uint64_t emu_xhuff_second_chances() { return g_xhuff_second; }
void emu_xhuff_decode_indexed(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                              const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid, uint32_t damage, uint32_t trusted)
{
    using namespace msc;
    std::vector<uint64_t> ioff(count + 1, 0);
    for (uint32_t i = 0; i < count; ++i) {
        const uint64_t by_out = xhi_max_segments(out_cap[i]), by_in = (uint64_t)in_len[i] / 8 + in_len[i] / 260 + 3;
        ioff[i + 1] = ioff[i] + std::min(by_out, by_in);
    }
    std::vector<XhSeg> index(ioff[count] + 1);
    std::vector<uint32_t> nseg(count + 1);
    emu::launch(xhuff_index_kernel, dim3e(grid), dim3e(32 * XHD_WARPS_PER_BLOCK), in, count, in_off, in_len, out_cap, index.data(), (const uint64_t*)ioff.data(),
                nseg.data(), out_len, status);
    if (damage) for (uint32_t i = 0; i < count; ++i) if (nseg[i] >= damage) index[ioff[i] + damage - 1].ip += 2;
    std::vector<uint4> tokbuf((size_t)grid * XHX_WARPS_PER_BLOCK * XHX_TOKBUF_WORDS / 4 + 4);
    emu::launch(xhuff_decode_indexed_kernel, dim3e(grid), dim3e(32 * XHX_WARPS_PER_BLOCK), in, out, count, in_off, in_len, out_off, out_cap,
                (const XhSeg*)index.data(), (const uint64_t*)ioff.data(), (const uint32_t*)nseg.data(), out_len, status, trusted ? 0 : 1, (uint32_t*)tokbuf.data());
    unsigned long long handled = 0;
    emu::launch(xhuff_decode_marked_kernel, dim3e(grid), dim3e(32 * XHD_WARPS_PER_BLOCK), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                (int)ST_INDEX, &handled);
    g_xhuff_second += handled;
}

// One multi-chunk Xpress-Huffman stream through the chunk-parallel path (kernels/xhuff_chunks.cuh): candidates -> one lane per
// candidate parses one chunk -> chain from offset 0 -> token execution (all chunks at once when none reaches back, else in order).
// The stitching mirrors msc_xhuff_stream_decode (csrc/mscomp_gpu.cu).  Returns 1 when the stream was decoded this way (*total
// bytes in out), 0 when the path declines; *info = candidates | chunks << 16 | (1 << 31 if the chunks ran independently).
static int g_chain_in_order = 0;
void emu_xhuff_stream_in_order(int on) { g_chain_in_order = on; }
static int g_chunk_lane_parse = 0;                                  // 1: the batch parser, one lane per candidate; 2: a warp per candidate, lane 0 walks; 0: a warp per candidate, speculative rounds (the default)
void emu_xhuff_stream_lane_parse(int on) { g_chunk_lane_parse = on; }
void emu_xhuff_chunk_warp_counts(uint64_t* out) { out[0] = msc::g_xhw_round_syms; out[1] = msc::g_xhw_serial_syms; msc::g_xhw_round_syms = msc::g_xhw_serial_syms = 0; }
uint32_t emu_xhuff_stream_decode(const uint8_t* in, uint32_t n, uint8_t* out, uint64_t cap, uint64_t* total, uint32_t* info)
{
    using namespace msc;
    *total = 0; *info = 0;
    if (n < XH_MIN_DATA + 2u) return 0;
    const uint32_t capc = n / XH_MIN_DATA + 2;
    std::vector<uint32_t> cand(capc + 2, 0);
    uint32_t ncand = 0;
    emu::launch(xhs_candidates_kernel, dim3e(3), dim3e(XHS_THREADS), in, n, cand.data(), &ncand, capc);
    if (ncand < 1 || ncand > capc) return 0;
    cand.resize(ncand);
    std::sort(cand.begin(), cand.end());
    *info = ncand;
    if (cand[0] != 0) return 0;
    const uint32_t count = ncand;
    std::vector<uint64_t> in_off(count), out_off(count);
    std::vector<uint32_t> in_len(count), out_cap(count, XH_CHUNK), out_len(count), ntok(count + 1), end_ip(count), reach(count);
    std::vector<int32_t> status(count);
    for (uint32_t k = 0; k < count; ++k) { in_off[k] = cand[k]; in_len[k] = n - cand[k]; out_off[k] = (uint64_t)k * XH_CHUNK; }
    std::vector<uint32_t> tok((size_t)count * XH_CHUNK + 16);
    if (g_chunk_lane_parse == 2)
        emu::launch(xhuff_parse_chunks_warp_serial_kernel, dim3e(2), dim3e(32 * XHW_WARPS_PER_BLOCK), in, tok.data(), count, (const uint64_t*)in_off.data(),
                    (const uint32_t*)in_len.data(), (const uint64_t*)out_off.data(), (const uint32_t*)out_cap.data(), out_len.data(), status.data(),
                    ntok.data(), end_ip.data(), reach.data());
    else if (g_chunk_lane_parse)
        emu::launch(xhuff_parse_chunks_kernel, dim3e(2), dim3e(32 * XHD_WARPS_PER_BLOCK), in, tok.data(), count, (const uint64_t*)in_off.data(),
                    (const uint32_t*)in_len.data(), (const uint64_t*)out_off.data(), (const uint32_t*)out_cap.data(), out_len.data(), status.data(),
                    ntok.data(), end_ip.data(), reach.data());
    else
        emu::launch(xhuff_parse_chunks_warp_kernel, dim3e(2), dim3e(32 * XHW_WARPS_PER_BLOCK), in, tok.data(), count, (const uint64_t*)in_off.data(),
                    (const uint32_t*)in_len.data(), (const uint64_t*)out_off.data(), (const uint32_t*)out_cap.data(), out_len.data(), status.data(),
                    ntok.data(), end_ip.data(), reach.data());
    std::vector<uint32_t> chain, pos;
    uint64_t produced = 0;
    uint32_t at = 0;
    bool ok = false, independent = true;
    for (;;) {
        const auto it = std::lower_bound(cand.begin(), cand.end(), at);
        if (it == cand.end() || *it != at) break;
        const uint32_t k = (uint32_t)(it - cand.begin());
        if (status[k] < 0 || reach[k] > produced) break;
        if (reach[k]) independent = false;
        chain.push_back(k); pos.push_back((uint32_t)produced);
        produced += out_len[k];
        if (status[k] == 0) { ok = true; break; }
        const uint64_t next = (uint64_t)at + end_ip[k];
        if (next == n) { ok = true; break; }
        if (next > n || n - next < (uint64_t)XH_MIN_DATA || out_len[k] == 0) break;
        at = (uint32_t)next;
    }
    if (!ok || produced > cap) return 0;
    {
        std::vector<uint64_t> slot(chain.size());
        std::vector<uint32_t> nt(chain.size());
        for (size_t j = 0; j < chain.size(); ++j) { slot[j] = (uint64_t)chain[j] * XH_CHUNK; nt[j] = ntok[chain[j]]; }
        if (independent && g_chain_in_order)                       // (the product takes this branch for streams of many chunks; few chunks go by pointer jumping)
            emu::launch(xhs_apply_chunks_kernel, dim3e(2), dim3e(32 * LZA_WARPS_PER_BLOCK), (const uint32_t*)tok.data(), (const uint64_t*)slot.data(),
                        (const uint32_t*)nt.data(), (const uint32_t*)pos.data(), (uint32_t)chain.size(), out);
        else if (g_chain_in_order)
            emu::launch(xhs_apply_chain_kernel, dim3e(1), dim3e(32 * LZB_WARPS), (const uint32_t*)tok.data(), (const uint64_t*)slot.data(),
                        (const uint32_t*)nt.data(), (const uint32_t*)pos.data(), (uint32_t)chain.size(), out);
        else {                                                     // pointer jumping (lz_jump.cuh), as msc_xhuff_stream_decode does it
            std::vector<uint32_t> word(produced + 1, 0xDEADBEEFu);
            uint32_t unknown = 0, rounds = 1;
            while ((1ull << rounds) < produced) ++rounds;
            emu::launch(lzj_expand_kernel, dim3e(2), dim3e(LZJ_THREADS), (const uint32_t*)tok.data(), (const uint64_t*)slot.data(), (const uint32_t*)nt.data(),
                        (const uint32_t*)pos.data(), (uint32_t)chain.size(), word.data());
            for (uint32_t r = 0; r < rounds; ++r) emu::launch(lzj_round_kernel, dim3e(3), dim3e(LZJ_THREADS), word.data(), (uint32_t)produced);
            emu::launch(lzj_store_kernel, dim3e(3), dim3e(LZJ_THREADS), (const uint32_t*)word.data(), out, (uint32_t)produced, &unknown);
            if (unknown) return 0;
        }
    }
    const bool packed = independent;
    *total = produced;
    *info = ncand | ((uint32_t)chain.size() << 16) | (packed ? 0x80000000u : 0u);
    return 1;
}

// two-pass plain Xpress encoder: match finding per 64 KiB segment (persistent blocks), then emission per buffer
void emu_xpress_encode2(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                        const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid,
                        uint32_t seg_bytes, void* index, uint32_t stride)
{
    using namespace msc;
    std::vector<uint32_t> seg_first(count + 2);
    unsigned long long totals[3] = {0, 0, 0};
    uint32_t work[4] = {0, 0, 0, 0};
    emu::launch(xpress_segments_kernel, dim3e(1), dim3e(1024), in_len, count, seg_first.data(), totals, work);
    const size_t tiles = (size_t)totals[0] * XPE_SEG_TILES;
    std::vector<uint32_t> lenoff(tiles * LZF_TILE + 1), visited(tiles * (LZF_TILE / 32) + 1), next_p(tiles + 1);
    if (totals[0])
        emu::launch(xpress_match_kernel, dim3e(grid), dim3e(XPE_THREADS), in, count, in_off, in_len, (const uint32_t*)seg_first.data(), work,
                    lenoff.data(), visited.data(), next_p.data());
    XpPre pre{seg_first.data(), lenoff.data(), visited.data(), next_p.data()};
    emu::launch(xpress_emit_kernel, dim3e(grid), dim3e(XPE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                index ? seg_bytes / (uint32_t)LZF_TILE : 0u, (XpressSeg*)index, stride, pre);
}

// the reference's parse: hash-chain links (one warp per 32 KiB piece), Find at every position, greedy parse + emission per buffer
void emu_xpress_encode_exact(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                             const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid,
                             uint32_t seg_bytes, void* index, uint32_t stride)
{
    using namespace msc;
    std::vector<uint32_t> seg_first(count + 2);
    unsigned long long totals[3] = {0, 0, 0};
    uint32_t work[4] = {0, 0, 0, 0};
    emu::launch(xpress_segments_kernel, dim3e(1), dim3e(1024), in_len, count, seg_first.data(), totals, work);
    const size_t npos = (size_t)totals[0] * LZX_SEG;
    std::vector<uint16_t> dist(npos + 1, 0xDEAD);
    std::vector<uint32_t> lenoff(npos + 1, LZX_UNKNOWN), visited0(npos / 32 + 1, 0xDEADBEEFu), exit0(npos / LZF_TILE + 1, 0xDEADBEEFu), exit_sub(npos / LZX_SUB + 1, 0xDEADBEEFu);
    if (totals[0]) {
        emu::launch(lz_prev_kernel<uint16_t>, dim3e(grid), dim3e(32), in, count, in_off, in_len, (const uint32_t*)seg_first.data(), work,
                    dist.data(), 8192u, LZX_PIECE);
        emu::launch(lz_find_kernel<LZX_FIND_MINB>, dim3e(grid), dim3e(LZX_FIND_THREADS), in, count, in_off, in_len, (const uint32_t*)seg_first.data(), work + 1,
                    (const uint16_t*)dist.data(), lenoff.data(), visited0.data(), exit_sub.data(), 8192u, 0u);
        emu::launch(lz_stitch_kernel, dim3e(grid), dim3e(128), in, count, in_off, in_len, (const uint32_t*)seg_first.data(), (const uint16_t*)dist.data(),
                    lenoff.data(), visited0.data(), (const uint32_t*)exit_sub.data(), exit0.data(), 8192u, 0u);
    }
    XpPre pre{seg_first.data(), lenoff.data(), visited0.data(), exit0.data(), dist.data()};
    emu::launch(xpress_emit_exact_kernel, dim3e(grid), dim3e(XPE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                index ? seg_bytes / (uint32_t)LZF_TILE : 0u, (XpressSeg*)index, stride, pre);
}

// few-stream variant: one warp parses a stream (a flag group of 32 items at a time), one block applies it
void emu_xpress_decode_warp(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                            const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid)
{
    std::vector<uint32_t> tok(emu_extent(count, out_off, out_cap) + 64), ntok(count + 1);
    emu::launch(msc::xpress_parse_warp_kernel, dim3e(grid), dim3e(32 * msc::XPD_WARPS_PER_BLOCK), in, tok.data(), count, in_off, in_len, out_off, out_cap, out_len, status, ntok.data());
    emu::launch(msc::lz_apply_block_kernel, dim3e(grid), dim3e(32 * msc::LZB_WARPS), (const uint32_t*)tok.data(), (const uint32_t*)ntok.data(), out_off, (const int32_t*)status, count, out);
}

// plain Xpress with the side-band restart index: encode writes it, decode parses one lane per segment
void emu_xpress_encode_indexed(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                               const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid,
                               uint32_t seg_bytes, void* index, uint32_t stride)
{
    emu::launch(msc::xpress_encode_kernel, dim3e(grid), dim3e(msc::XPE_THREADS), in, out, count, in_off, in_len, out_off, out_cap, out_len, status,
                seg_bytes / (uint32_t)msc::LZF_TILE, (msc::XpressSeg*)index, stride);
}

void emu_xpress_decode_indexed(const uint8_t* in, uint8_t* out, uint32_t count, const uint64_t* in_off, const uint32_t* in_len,
                               const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t grid,
                               const void* index, uint32_t stride, uint32_t block_apply)
{
    const size_t lanes = (size_t)count * stride;
    std::vector<uint32_t> tok(emu_extent(count, out_off, out_cap) + 16), seg_ntok(lanes + 1), seg_end(lanes + 1);
    std::vector<int32_t> seg_status(lanes + 1);
    emu::launch(msc::xpress_parse_indexed_kernel, dim3e(grid), dim3e(32 * msc::XPD_WARPS_PER_BLOCK), in, tok.data(), count, in_off, in_len, out_off, out_cap,
                (const msc::XpressSeg*)index, stride, seg_ntok.data(), seg_status.data(), seg_end.data());
    if (block_apply)
        emu::launch(msc::xpress_apply_indexed_block_kernel, dim3e(grid), dim3e(32 * msc::LZB_WARPS), (const uint32_t*)tok.data(), (const uint32_t*)seg_ntok.data(),
                    (const int32_t*)seg_status.data(), (const uint32_t*)seg_end.data(), (const msc::XpressSeg*)index, stride, out_off, count, out, out_len, status);
    else
        emu::launch(msc::xpress_apply_indexed_kernel, dim3e(grid), dim3e(32 * msc::LZA_WARPS_PER_BLOCK), (const uint32_t*)tok.data(), (const uint32_t*)seg_ntok.data(),
                    (const int32_t*)seg_status.data(), (const uint32_t*)seg_end.data(), (const msc::XpressSeg*)index, stride, out_off, count, out, out_len, status);
}

// One plain Xpress stream, flag groups by speculation (kernels/xpress_spec.cuh), as msc_xpress_stream_decode does it.
// returns 1 = decoded (*total bytes), 0 = left to the serial decoder; info = {blocks, groups, restart entries, stage that gave up}
uint32_t emu_xpress_stream_decode(const uint8_t* in, uint32_t n, uint8_t* out, uint64_t cap, uint64_t* total, uint32_t* info)
{
    using namespace msc;
    *total = 0; info[0] = info[1] = info[2] = info[3] = 0;
    if (n < 64) return 0;
    const uint32_t npos = n - 3, nblk = (npos + XPS_BLK - 1) / XPS_BLK;
    info[0] = nblk;
    std::vector<unsigned long long> rec((size_t)npos * 3 + 8), count(nblk + 1), base(nblk + 1);
    std::vector<uint32_t> table((size_t)nblk * XPS_NENT + 8), entry(nblk + 1), entry_half(nblk + 1);
    XpsSummary sum;
    emu::launch(xps_groups_kernel, dim3e(3), dim3e(256), in, n, rec.data());
    emu::launch(xps_blocks_kernel, dim3e(2), dim3e(XPS_THREADS), (const unsigned long long*)rec.data(), n, nblk, table.data());
    emu::launch(xps_entries_kernel, dim3e(1), dim3e(32), (const uint32_t*)table.data(), nblk, entry.data(), entry_half.data(), &sum);
    emu::launch(xps_walk_kernel<false>, dim3e(3), dim3e(32), in, n, (const unsigned long long*)rec.data(), (const uint32_t*)entry.data(),
                (const uint32_t*)entry_half.data(), (const XpsSummary*)&sum, count.data(), (const unsigned long long*)base.data(), (XpressSeg*)nullptr);
    emu::launch(xps_scan_kernel, dim3e(1), dim3e(32), (const unsigned long long*)count.data(), base.data(), &sum, (unsigned long long)cap);
    info[1] = sum.groups; info[2] = sum.nseg;
    if (!sum.ok || !sum.nseg || sum.bytes > cap || sum.bytes > 0x7FFFFFFFull || !sum.bytes) { info[3] = 1; return 0; }
    const uint32_t nseg = sum.nseg, produced = (uint32_t)sum.bytes;
    std::vector<XpressSeg> index(nseg + 1);
    std::vector<uint32_t> tok((size_t)produced + 64), seg_ntok(nseg + 1), seg_end(nseg + 1), ntok(nseg + 1), pos(nseg + 1);
    std::vector<int32_t> seg_status(nseg + 1);
    std::vector<unsigned long long> slot(nseg + 1);
    uint32_t result[4] = {0, 0, 0, 0};
    const uint64_t in_off = 0, out_off = 0;
    const uint32_t in_len = n, out_cap = produced;
    emu::launch(xps_walk_kernel<true>, dim3e(3), dim3e(32), in, n, (const unsigned long long*)rec.data(), (const uint32_t*)entry.data(),
                (const uint32_t*)entry_half.data(), (const XpsSummary*)&sum, count.data(), (const unsigned long long*)base.data(), index.data());
    emu::launch(xpress_parse_indexed_kernel, dim3e(2), dim3e(32 * XPD_WARPS_PER_BLOCK), in, tok.data(), 1u, &in_off, &in_len, &out_off, &out_cap,
                (const XpressSeg*)index.data(), nseg, seg_ntok.data(), seg_status.data(), seg_end.data());
    emu::launch(xps_chain_kernel, dim3e(2), dim3e(256), (const XpressSeg*)index.data(), nseg, (const uint32_t*)seg_ntok.data(), (const int32_t*)seg_status.data(),
                (const uint32_t*)seg_end.data(), slot.data(), ntok.data(), pos.data(), result);
    if (result[0] != 0 || result[1] != produced) { info[3] = 2; return 0; }
    std::vector<uint32_t> word((size_t)produced + 1, 0xDEADBEEFu);
    uint32_t unknown = 0, rounds = 1;
    while ((1ull << rounds) < produced) ++rounds;
    emu::launch(lzj_expand_kernel, dim3e(2), dim3e(LZJ_THREADS), (const uint32_t*)tok.data(), (const uint64_t*)slot.data(), (const uint32_t*)ntok.data(),
                (const uint32_t*)pos.data(), nseg, word.data());
    for (uint32_t r = 0; r < rounds; ++r) emu::launch(lzj_round_kernel, dim3e(3), dim3e(LZJ_THREADS), word.data(), produced);
    emu::launch(lzj_store_kernel, dim3e(3), dim3e(LZJ_THREADS), (const uint32_t*)word.data(), out, produced, &unknown);
    if (unknown) { info[3] = 3; return 0; }
    *total = produced;
    return 1;
}

// Chunk index of an LZNT1 stream, built range by range (range_segs segments at a time) like the pipelined host wrapper does.
// state_out = {done, tail, ip, nchunks}; returns 0, or 1 when more than cap chunks were found.
int emu_lznt1_index(const uint8_t* in, uint64_t in_len, uint32_t S, uint32_t range_segs, uint64_t* offs, uint64_t cap, uint64_t* state_out)
{
    using namespace msc;
    const uint64_t ns64 = in_len ? (in_len + S - 1) / S : 1;
    const uint32_t NS = (uint32_t)ns64;
    Lznt1IndexState st;
    memset(&st, 0, sizeof(st));
    for (uint32_t seg0 = 0; seg0 < NS; seg0 += range_segs) {
        const uint32_t nseg = std::min(range_segs, NS - seg0), ng = (nseg + LZI_GROUP - 1) / LZI_GROUP;
        std::vector<uint16_t> T1((size_t)nseg * LZI_STRIDE), T2((size_t)ng * LZI_STRIDE), seg_entry(nseg), grp_entry(ng);
        std::vector<uint32_t> seg_base(nseg);
        emu::launch(lznt1_index_tables_kernel, dim3e(std::min(nseg, 3u)), dim3e(LZI_THREADS), in, in_len, seg0, nseg, S, T1.data());
        emu::launch(lznt1_index_compose_kernel, dim3e(ng), dim3e(LZI_THREADS), (const uint16_t*)T1.data(), nseg, T2.data());
        emu::launch(lznt1_index_entries_kernel, dim3e(1), dim3e(256), (const uint16_t*)T1.data(), (const uint16_t*)T2.data(), nseg, &st, seg_entry.data(), grp_entry.data());
        emu::launch(lznt1_index_count_kernel, dim3e(1), dim3e(LZI_THREADS), in, in_len, seg0, nseg, S, (const uint16_t*)seg_entry.data(), seg_base.data(), &st, (Lznt1IndexState*)nullptr);
        if (st.nchunks > cap) return 1;
        emu::launch(lznt1_index_write_kernel, dim3e(2), dim3e(256), in, in_len, seg0, nseg, S, (const uint16_t*)seg_entry.data(), (const uint32_t*)seg_base.data(),
                    offs + (st.nchunks - st.range_chunks));
    }
    state_out[0] = st.done; state_out[1] = st.tail; state_out[2] = st.ip; state_out[3] = st.nchunks;
    return 0;
}

}
