// common.cuh -- shared device helpers for the sm_100a codec kernels.
// The same sources also compile with g++ under tests/emu/simt_emu.h (MSC_EMU) for the CPU test-suite.
#pragma once
#ifndef MSC_EMU
#include <cuda_runtime.h>
#include <stdint.h>
#else
#include <string.h>
#endif

namespace msc {

// Status codes shared with the C ABI (include/mscomp/general.h).
enum : int { ST_OK = 0, ST_ARG = -2, ST_DATA = -3, ST_MEM = -4, ST_BUF = -5 };

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31); }

// 16-byte streaming load (read-once data: do not allocate in L1).
__device__ __forceinline__ uint4 ld_stream16(const void* p)
{
#ifndef MSC_EMU
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return *reinterpret_cast<const uint4*>(p);
#endif
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v)
{
#ifndef MSC_EMU
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    *reinterpret_cast<uint4*>(p) = v;
#endif
}

// One byte of global memory that another warp of this block may have written a moment ago: bypass L1.
__device__ __forceinline__ uint32_t ld_byte_cg(const uint8_t* p)
{
#ifndef MSC_EMU
    uint32_t r;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
#else
    return *p;
#endif
}
// Shared-memory words other warps update while this one polls them: atomics on both sides (a deliberate signal between
// warps, and the form compute-sanitizer's racecheck understands as such).
__device__ __forceinline__ uint32_t ld_shared_signal(uint32_t* p) { return atomicAdd(p, 0u); }
__device__ __forceinline__ void st_shared_signal(uint32_t* p, uint32_t v) { atomicExch(p, v); }
__device__ __forceinline__ void spin_pause()
{
#ifndef MSC_EMU
    __nanosleep(40);
#else
    emu::yield();
#endif
}

// ---- bulk asynchronous copies (TMA, 1-D) + mbarrier ---------------------------------------------------------------
// cp.async.bulk moves a 16-byte-aligned, 16-byte-multiple range between global and shared memory without passing
// through registers; a load signals an mbarrier with the byte count, a store is tracked by a bulk group.
// Under the CPU emulator the copies are plain memcpy and the barrier is a no-op.
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
#ifndef MSC_EMU
    return (uint32_t)__cvta_generic_to_shared(p);
#else
    return (uint32_t)(uintptr_t)p;
#endif
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
#ifndef MSC_EMU
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    *bar = 0; (void)count;
#endif
}
// writes made through the generic proxy (ordinary stores) become visible to / ordered before later async-proxy accesses
__device__ __forceinline__ void fence_async_smem()
{
#ifndef MSC_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// one thread: announce `bytes` on the barrier and start the global -> shared copy that will deliver them
__device__ __forceinline__ void bulk_load_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
#ifndef MSC_EMU
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
#else
    memcpy(dst_smem, src_gmem, bytes); (void)bar;
#endif
}
// every thread that reads the data: wait for the phase with the given parity to complete
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
#ifndef MSC_EMU
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
#else
    (void)bar; (void)parity;
    __syncwarp();                       // emulator: the copy is a memcpy made by one lane; the waiters meet it here (warp-scope use only)
#endif
}
// one thread: shared -> global copy as a bulk group; wait_read returns once the shared source may be overwritten
__device__ __forceinline__ void bulk_store_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
#ifndef MSC_EMU
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#else
    memcpy(dst_gmem, src_smem, bytes);
#endif
}
__device__ __forceinline__ void bulk_store_wait_read()
{
#ifndef MSC_EMU
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

// One byte of shared memory by its 32-bit shared-window address `saddr` (base / base_saddr = generic pointer and shared
// address of one reference byte: the emulator, where a 32-bit address is not enough, reads base[saddr - base_saddr]).
__device__ __forceinline__ uint32_t lds_u8(const uint8_t* base, uint32_t base_saddr, uint32_t saddr)
{
#ifndef MSC_EMU
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    (void)base; (void)base_saddr;
    return v;
#else
    return base[saddr - base_saddr];
#endif
}

// Inclusive warp prefix sum.
__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Warp-cooperative copy of `n` bytes global/shared -> global, arbitrary alignment of both sides.
// Uses 16-byte stores on the destination-aligned body and builds each vector from byte loads only at the
// ragged ends; the body loads are 4-byte aligned-down words recombined with a funnel shift.
__device__ __forceinline__ void warp_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n)
{
    const int lane = lane_id();
    // head: bring dst to 16-byte alignment
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t nvec = n >> 4;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 3u);
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src - sh);
    for (uint32_t v = lane; v < nvec; v += 32) {
        const uint32_t* p = s32 + 4 * v;
        uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
        uint4 o;
        if (sh == 0) { o = make_uint4(w0, w1, w2, w3); }
        else {
            uint32_t w4 = p[4];
            o.x = __funnelshift_r(w0, w1, 8 * sh); o.y = __funnelshift_r(w1, w2, 8 * sh);
            o.z = __funnelshift_r(w2, w3, 8 * sh); o.w = __funnelshift_r(w3, w4, 8 * sh);
        }
        *reinterpret_cast<uint4*>(dst + 16 * v) = o;
    }
    const uint32_t done = nvec << 4;
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

} // namespace msc

// Dynamic shared memory declaration usable by both builds.
#ifndef MSC_EMU
#define MSC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define MSC_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#endif
