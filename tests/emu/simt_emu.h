/*
 * simt_emu.h -- TEST INFRASTRUCTURE.  A minimal single-process SIMT emulator so that the CUDA kernel sources
 * under ms-compress_b200/csrc/kernels/ can ALSO be compiled by g++ and executed on the CPU by the `-m "not gpu"`
 * test-suite (with -fsanitize=address,undefined): it finds indexing/logic errors in the warp-cooperative code
 * without spending GPU time.  It is never part of the product library (which has no CPU path).
 *
 * Model: one block at a time; every CUDA thread is a stackful fiber (hand-rolled x86-64 context switch);
 * fibers run until they reach a collective (__syncwarp / __shfl_* / __ballot / __syncthreads ...), where they
 * rendezvous.  Only full-mask warp collectives at convergent points are supported -- which is also the only
 * style the kernels use.  Threads that have returned count as arrived.
 */
#pragma once
#ifndef __x86_64__
#error "simt_emu.h needs x86-64"
#endif
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

#define MSC_EMU 1

namespace emu {

struct dim3e { unsigned x = 1, y = 1, z = 1; dim3e() {} dim3e(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

struct Fiber { void* sp = nullptr; char* stack = nullptr; dim3e tid; bool done = false; int warp = 0, lane = 0; };
struct Warp { int arrived = 0, alive = 0; unsigned gen = 0; uint64_t slot[2][32]; };
struct Block {
    std::vector<Fiber> fibers; std::vector<Warp> warps;
    int nthreads = 0, cur = 0, alive = 0, blk_arrived = 0; unsigned blk_gen = 0;
    dim3e bid, bdim, gdim; void* sched_sp = nullptr; std::function<void()> body;
};
inline Block*& blk() { static Block* b = nullptr; return b; }
inline unsigned char* dyn_smem() { alignas(128) static unsigned char buf[232448]; return buf; }

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    " pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
    " movq %rsp, (%rdi)\n movq %rsi, %rsp\n"
    " popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n ret\n"
    ".size emu_switch,.-emu_switch\n");

inline Fiber* cur() { Block* b = blk(); return &b->fibers[b->cur]; }
inline void yield() { Block* b = blk(); emu_switch(&b->fibers[b->cur].sp, b->sched_sp); }

inline void trampoline()
{
    Block* b = blk();
    b->body();
    Fiber* f = &b->fibers[b->cur];
    f->done = true; b->alive--; b->warps[f->warp].alive--;
    for (;;) yield();
}

constexpr size_t kStack = 256 * 1024;

template <class K, class... A>
void launch(K kernel, dim3e grid, dim3e block, A... args)
{
    const int nt = (int)(block.x * block.y * block.z);
    std::vector<char*> stacks(nt);
    for (int t = 0; t < nt; ++t) stacks[t] = (char*)aligned_alloc(64, kStack);
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block b;
        b.nthreads = nt; b.alive = nt; b.bid = dim3e(bx, by, bz); b.bdim = block; b.gdim = grid;
        b.fibers.resize(nt); b.warps.resize((nt + 31) / 32);
        for (int t = 0; t < nt; ++t) {
            Fiber& f = b.fibers[t];
            f.tid = dim3e(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            f.warp = t / 32; f.lane = t % 32; f.stack = stacks[t];
            b.warps[f.warp].alive++;
            uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
            void** sp = (void**)top;
            *--sp = nullptr;                       /* fake return address of the trampoline */
            *--sp = (void*)&trampoline;            /* popped by `ret` in emu_switch */
            for (int i = 0; i < 6; ++i) *--sp = nullptr;
            f.sp = (void*)sp;
        }
        b.body = [&]() { kernel(args...); };
        blk() = &b;
        int idle_rounds = 0;
        while (b.alive > 0) {
            bool progressed = false;
            for (int t = 0; t < nt && b.alive > 0; ++t) {
                if (b.fibers[t].done) continue;
                b.cur = t;
                int before = b.alive;
                emu_switch(&b.sched_sp, b.fibers[t].sp);
                if (b.alive != before) progressed = true;
            }
            /* deadlock guard: if nothing finishes for a very long time, abort loudly */
            idle_rounds = progressed ? 0 : idle_rounds + 1;
            if (idle_rounds > 200000000) { fprintf(stderr, "simt_emu: deadlock suspected\n"); abort(); }
        }
        blk() = nullptr;
    }
    for (int t = 0; t < nt; ++t) free(stacks[t]);
}

inline void warp_rendezvous()
{
    Block* b = blk(); Fiber* f = cur(); Warp& w = b->warps[f->warp];
    const unsigned g = w.gen;
    w.arrived++;
    for (;;) {
        if (w.gen != g) return;
        if (w.arrived >= w.alive) { w.arrived = 0; w.gen++; return; }
        yield();
    }
}
inline void block_rendezvous()
{
    Block* b = blk();
    const unsigned g = b->blk_gen;
    b->blk_arrived++;
    for (;;) {
        if (b->blk_gen != g) return;
        if (b->blk_arrived >= b->alive) { b->blk_arrived = 0; b->blk_gen++; return; }
        yield();
    }
}
inline uint64_t* exchange(uint64_t v)
{   /* publish v, rendezvous, return the 32-slot table of this collective */
    Block* b = blk(); Fiber* f = cur(); Warp& w = b->warps[f->warp];
    const unsigned p = w.gen & 1;
    w.slot[p][f->lane] = v;
    warp_rendezvous();
    return w.slot[p];
}
inline int warp_lanes() { Block* b = blk(); Fiber* f = cur(); int base = f->warp * 32; return std::min(32, b->nthreads - base); }

template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "shuffle type"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

} // namespace emu

/* ---- CUDA surface ---------------------------------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define threadIdx (emu::cur()->tid)
#define blockIdx (emu::blk()->bid)
#define blockDim (emu::blk()->bdim)
#define gridDim (emu::blk()->gdim)
#define warpSize 32

struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

#define EMU_FULL(mask) assert((mask) == 0xffffffffu)

inline void __syncwarp(unsigned mask = 0xffffffffu) { EMU_FULL(mask); emu::warp_rendezvous(); }
inline void __syncthreads() { emu::block_rendezvous(); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{ EMU_FULL(mask); (void)width; int me = emu::cur()->lane; uint64_t* s = emu::exchange(emu::to_bits(v)); int l = src & 31; return emu::from_bits<T>(l < emu::warp_lanes() ? s[l] : s[me]); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{ EMU_FULL(mask); (void)width; int me = emu::cur()->lane; uint64_t* s = emu::exchange(emu::to_bits(v)); int l = me - (int)d; return emu::from_bits<T>(l >= 0 ? s[l] : s[me]); }
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{ EMU_FULL(mask); (void)width; int me = emu::cur()->lane; uint64_t* s = emu::exchange(emu::to_bits(v)); int l = me + (int)d; return emu::from_bits<T>(l < emu::warp_lanes() ? s[l] : s[me]); }
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{ EMU_FULL(mask); (void)width; int me = emu::cur()->lane; uint64_t* s = emu::exchange(emu::to_bits(v)); int l = me ^ x; return emu::from_bits<T>(l < emu::warp_lanes() ? s[l] : s[me]); }

inline unsigned __ballot_sync(unsigned mask, int pred)
{ EMU_FULL(mask); uint64_t* s = emu::exchange(pred ? 1 : 0); unsigned r = 0; int n = emu::warp_lanes();
  emu::Block* b = emu::blk(); int base = emu::cur()->warp * 32;
  for (int l = 0; l < n; ++l) if (!b->fibers[base + l].done && s[l]) r |= 1u << l; return r; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
inline unsigned __activemask() { return 0xffffffffu; }

template <class T, class F> inline T emu_reduce(unsigned mask, T v, F f)
{ EMU_FULL(mask); uint64_t* s = emu::exchange(emu::to_bits(v)); int n = emu::warp_lanes(); T r = v; emu::Block* b = emu::blk(); int base = emu::cur()->warp * 32;
  for (int l = 0; l < n; ++l) if (!b->fibers[base + l].done) r = f(r, emu::from_bits<T>(s[l])); return r; }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a + b; }) - v; }
inline int __reduce_min_sync(unsigned m, int v) { return emu_reduce(m, v, [](int a, int b) { return a < b ? a : b; }); }
inline int __reduce_max_sync(unsigned m, int v) { return emu_reduce(m, v, [](int a, int b) { return a > b ? a : b; }); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a | b; }); }
inline unsigned __match_any_sync(unsigned mask, unsigned v)
{ EMU_FULL(mask); uint64_t* s = emu::exchange(v); unsigned r = 0; int n = emu::warp_lanes(); emu::Block* b = emu::blk(); int base = emu::cur()->warp * 32;
  for (int l = 0; l < n; ++l) if (!b->fibers[base + l].done && (unsigned)s[l] == v) r |= 1u << l; return r; }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{ uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0; for (int i = 0; i < 4; ++i) { unsigned k = (sel >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * k)) & 0xFF) << (8 * i); } return r; }
template <class T> inline T __ldg(const T* p) { return *p; }
/* min/max: kernels use the two-argument integer forms */
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }

template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
