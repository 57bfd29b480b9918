// stream_host.cpp -- TEST INFRASTRUCTURE.  The product's LZNT1 stream state machine (ms-compress_b200/csrc/stream_logic.h)
// instantiated with the ORACLE's chunk functions instead of the GPU batch codec, so that the `-m "not gpu"` suite can drive the
// very same host logic against the compiled reference, call by call.  Never part of the product library.
#include "../../ms-compress_b200/csrc/stream_logic.h"
#include "../../ms-compress_b200/csrc/lznt1_host.h"
#define MSC_HOST_ONLY 1
#include "../../ms-compress_b200/csrc/kernels/xpress_stream.cuh"
#include <new>
extern "C" {
#include "../../oracle/oracle.h"
}

using namespace msc_stream;

namespace {
struct OracleCodec {
    static MSCompStatus encode(const uint8_t* in, size_t n, std::vector<uint8_t>& comp, std::vector<uint64_t>& offs)
    {
        comp.clear(); offs.clear();
        for (size_t p = 0; p < n; p += CHUNK) {
            const size_t len = n - p < CHUNK ? n - p : CHUNK;
            uint8_t buf[CHUNK + 64];
            size_t sz = sizeof(buf);
            if (orc_lznt1_compress(in + p, len, buf, &sz) != 0) return MSCOMP_MEM_ERROR;
            offs.push_back(comp.size());
            comp.insert(comp.end(), buf, buf + sz);
        }
        offs.push_back(comp.size());
        return MSCOMP_OK;
    }
    static MSCompStatus decode(const uint8_t* in, size_t n, const std::vector<uint64_t>& offs, std::vector<uint8_t>& out,
                               std::vector<uint32_t>& sizes, std::vector<int32_t>& status)
    {
        const size_t k = offs.size();
        out.assign(k * CHUNK, 0); sizes.assign(k, 0); status.assign(k, 0);
        for (size_t i = 0; i < k; ++i) {
            const uint8_t* c = in + offs[i];
            const unsigned h = (unsigned)c[0] | ((unsigned)c[1] << 8);
            const size_t payload = (h & 0x0FFF) + 1;
            if (offs[i] + 2 + payload > n) { status[i] = -5; continue; }
            if (h & 0x8000) {
                size_t osz = 0;
                status[i] = orc_lznt1_decompress_chunk(c + 2, payload, out.data() + i * CHUNK, CHUNK, &osz);
                if (status[i] == -5) status[i] = -3;                 // (the stream layer reports both as a data error)
                sizes[i] = (uint32_t)osz;
            } else {
                memcpy(out.data() + i * CHUNK, c + 2, payload);
                sizes[i] = (uint32_t)payload;
            }
        }
        return MSCOMP_OK;
    }
};
} // namespace

namespace {
// the product's xpress_inflate host logic over the same step function, running in host memory
struct HostXpressStream {
    static void* create() { msc::XpStreamState* s = new (std::nothrow) msc::XpStreamState; if (s) msc::xps_init(s); return s; }
    static void destroy(void* h) { delete (msc::XpStreamState*)h; }
    static MSCompStatus step(void* h, const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t n_out, uint32_t* used, uint32_t* made, int* status)
    {
        msc::XpStreamResult r{0, 0, 0, 0};
        msc::xps_step((msc::XpStreamState*)h, in, n_in, out, n_out, &r);
        *used = r.in_used; *made = r.out_made; *status = r.status;
        return MSCOMP_OK;
    }
    static void tail(void* h, uint32_t* flags, uint32_t* flagged, uint32_t* kept)
    { const msc::XpStreamState* s = (const msc::XpStreamState*)h; *flags = s->flags; *flagged = s->flagged; *kept = s->stash_n; }
};
} // namespace

extern "C" {
MSCompStatus hs_xpress_inflate_init(mscomp_stream* s) { return xpress_inflate_init_impl<HostXpressStream>(s); }
// the same with the reference's bookkeeping slip reproduced (differential runs against the compiled reference)
MSCompStatus hs_xpressq_inflate_init(mscomp_stream* s)
{
    const MSCompStatus st = xpress_inflate_init_impl<HostXpressStream>(s);
    if (st == MSCOMP_OK) ((msc::XpStreamState*)s->state)->upstream_quirk = 1;
    return st;
}
MSCompStatus hs_xpressq_inflate(mscomp_stream* s) { return xpress_inflate_impl<HostXpressStream>(s); }
MSCompStatus hs_xpressq_inflate_end(mscomp_stream* s) { return xpress_inflate_end_impl<HostXpressStream>(s); }
MSCompStatus hs_xpress_inflate(mscomp_stream* s) { return xpress_inflate_impl<HostXpressStream>(s); }
MSCompStatus hs_xpress_inflate_end(mscomp_stream* s) { return xpress_inflate_end_impl<HostXpressStream>(s); }
MSCompStatus hs_xpress_deflate_init(mscomp_stream* s) { return xpress_deflate_init_impl(s); }
MSCompStatus hs_xpress_deflate(mscomp_stream* s, MSCompFlush f) { return xpress_deflate_impl(s, f); }
MSCompStatus hs_xpress_deflate_end(mscomp_stream* s) { return xpress_deflate_end_impl(s); }
MSCompStatus hs_lznt1_deflate_init(mscomp_stream* s) { return lznt1_deflate_init_impl(s); }
MSCompStatus hs_lznt1_deflate(mscomp_stream* s, MSCompFlush flush) { return lznt1_deflate_impl<OracleCodec>(s, flush); }
MSCompStatus hs_lznt1_deflate_end(mscomp_stream* s) { return lznt1_deflate_end_impl(s); }
MSCompStatus hs_lznt1_inflate_init(mscomp_stream* s) { return lznt1_inflate_init_impl(s); }
MSCompStatus hs_lznt1_inflate(mscomp_stream* s) { return lznt1_inflate_impl<OracleCodec>(s); }
MSCompStatus hs_lznt1_inflate_end(mscomp_stream* s) { return lznt1_inflate_end_impl(s); }
// The simple path of the product's one-shot lznt1_decompress (mscomp_shim.cu) with the oracle as the chunk decoder: header walk,
// per-chunk decode into 4096-byte slots, replay of the reference's driver for the status, compaction of the decoded bytes.
MSCompStatus hs_lznt1_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t* out_len)
{
    using namespace msc_host;
    const size_t cap = *out_len;
    std::vector<uint64_t> offs;
    size_t ip = 0;
    const Lznt1Tail tail = lznt1_walk(in, in_len, offs, &ip);
    std::vector<uint8_t> slots; std::vector<uint32_t> sizes; std::vector<int32_t> status;
    if (!offs.empty()) OracleCodec::decode(in, ip, offs, slots, sizes, status);
    size_t total = 0;
    const MSCompStatus st = lznt1_replay(offs.size(), cap, tail, [&](size_t c) { return (size_t)sizes[c]; }, [&](size_t c) { return status[c]; }, &total);
    if (st != MSCOMP_OK) return st;
    size_t op = 0;
    for (size_t c = 0; c < offs.size(); ++c) { memcpy(out + op, slots.data() + c * CHUNK, sizes[c]); op += sizes[c]; }
    *out_len = total;
    return MSCOMP_OK;
}
MSCompStatus hs_copy_init(mscomp_stream* s) { return copy_init(s); }
MSCompStatus hs_copy_deflate(mscomp_stream* s, MSCompFlush flush) { return copy_deflate(s, flush); }
MSCompStatus hs_copy_inflate(mscomp_stream* s) { return copy_inflate(s); }
MSCompStatus hs_copy_end(mscomp_stream* s) { return copy_end(s); }
}
