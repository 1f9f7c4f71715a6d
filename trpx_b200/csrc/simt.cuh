// simt.cuh -- the thin layer the kernels are written against.
//
// Product build (nvcc, sm_100a): every wrapper below is the CUDA intrinsic or the inline PTX it
// names (mbarrier, cp.async.bulk = TMA bulk copies, relaxed/acquire global accesses).
//
// TEST build (-DTRPX_EMU, plain g++): the same kernel sources are compiled for the host and run by
// tests/emu/emu.hpp, a cooperative-fiber SIMT emulator (one fiber per CUDA thread, blocks run one
// after another).  That build exists only so `pytest -m "not gpu"` can check the kernels' index
// arithmetic against the oracle without a GPU; it is never linked into libtrpx_b200.so and is not a
// fallback: the product library has no host execution path.
#pragma once
#include <tuple>

#include <stddef.h>
#include <stdint.h>

#ifdef TRPX_EMU
#include "emu.hpp"
#else
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#endif

namespace trpx {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

#ifndef TRPX_EMU
// ----------------------------------------------------------------------------- CUDA backend
#define TRPX_DEVICE __device__ __forceinline__
#define TRPX_DEVICE_NOINLINE __device__ __noinline__
#define TRPX_GRID_CONSTANT __grid_constant__
#define TRPX_HD __host__ __device__ __forceinline__
#define TRPX_KERNEL __global__
#define TRPX_SHARED __shared__
#define TRPX_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define TRPX_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)

TRPX_DEVICE u32 tid() { return threadIdx.x; }
TRPX_DEVICE u32 bid() { return blockIdx.x; }
TRPX_DEVICE u32 nthreads() { return blockDim.x; }
TRPX_DEVICE u32 nblocks() { return gridDim.x; }
TRPX_DEVICE void sync_block() { __syncthreads(); }
TRPX_DEVICE void sync_warp() { __syncwarp(); }
// named barrier `id` (1..15) over `n` threads (a multiple of 32): lets the worker warps of a
// warp-specialised CTA synchronise among themselves without the resolver warps
TRPX_DEVICE void bar_sync(u32 id, u32 n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// producer side of a named barrier: counts the calling warp(s) in without waiting; the consumer blocks in
// bar_sync() -- a hardware wait, no spin loop
TRPX_DEVICE void bar_arrive(u32 id, u32 n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
TRPX_DEVICE void spin_hint() { __nanosleep(32); }
TRPX_DEVICE void trap() { __trap(); }

TRPX_DEVICE u32 shfl(u32 v, int src) { return __shfl_sync(0xffffffffu, v, src); }
TRPX_DEVICE u64 shfl(u64 v, int src) { return __shfl_sync(0xffffffffu, v, src); }
TRPX_DEVICE u32 shfl_up(u32 v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
TRPX_DEVICE u64 shfl_up(u64 v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
TRPX_DEVICE u32 shfl_down(u32 v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
TRPX_DEVICE u32 shfl_xor(u32 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
TRPX_DEVICE u32 ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
TRPX_DEVICE bool all_lanes(bool p) { return __all_sync(0xffffffffu, p); }
TRPX_DEVICE bool any_lane(bool p) { return __any_sync(0xffffffffu, p); }
TRPX_DEVICE u32 warp_max(u32 v) { return __reduce_max_sync(0xffffffffu, v); }
TRPX_DEVICE u32 warp_min_u32(u32 v) { return __reduce_min_sync(0xffffffffu, v); }
TRPX_DEVICE u32 warp_or(u32 v) { return __reduce_or_sync(0xffffffffu, v); }
TRPX_DEVICE u32 warp_add(u32 v) { return __reduce_add_sync(0xffffffffu, v); }

TRPX_DEVICE int clz32(u32 x) { return __clz((int)x); }
TRPX_DEVICE int clz64(u64 x) { return __clzll((long long)x); }
TRPX_DEVICE int ffs32(u32 x) { return __ffs((int)x); }
TRPX_DEVICE int ffs64(u64 x) { return __ffsll((long long)x); }
TRPX_DEVICE int popc32(u32 x) { return __popc(x); }
TRPX_DEVICE u32 funnel_r(u32 lo, u32 hi, u32 sh) { return __funnelshift_r(lo, hi, sh); }   // sh & 31
TRPX_DEVICE u32 funnel_l(u32 lo, u32 hi, u32 sh) { return __funnelshift_l(lo, hi, sh); }   // (hi:lo << sh) >> 32
TRPX_DEVICE u32 funnel_rc(u32 lo, u32 hi, u32 sh) { return __funnelshift_rc(lo, hi, sh); } // (hi:lo >> min(sh, 32)), low word
TRPX_DEVICE u32 vabs2(u32 x) { return __vabs2(x); }   // |.| per signed 16-bit half (wraps for -32768)
TRPX_DEVICE u32 vabs4(u32 x) { return __vabs4(x); }   // |.| per signed byte

TRPX_DEVICE u32 atomic_add(u32* p, u32 v) { return atomicAdd(p, v); }
TRPX_DEVICE u64 atomic_add(u64* p, u64 v) { return atomicAdd(p, v); }
TRPX_DEVICE u32 atomic_or(u32* p, u32 v) { return atomicOr(p, v); }
TRPX_DEVICE u32 atomic_max(u32* p, u32 v) { return atomicMax(p, v); }

// 64-bit descriptor words: value and status travel in ONE word, so relaxed accesses suffice
TRPX_DEVICE u64 ld_relaxed(const u64* p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
TRPX_DEVICE void st_relaxed(u64* p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
TRPX_DEVICE u32 ld_relaxed(const u32* p)
{
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// A global load the compiler cannot reason about (it would otherwise turn loads from CTA-uniform addresses into
// uniform-register values at once -- and wait for them -- which defeats a software prefetch).
TRPX_DEVICE u64 ldg_u64_opaque(const u64* p)
{
    u64 v;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// streaming (read-once / write-once) global accesses: keep them out of L1
TRPX_DEVICE u32 ld_stream(const u32* p) { return __ldcs(p); }
TRPX_DEVICE void st_stream(u32* p, u32 v) { __stcs(p, v); }
TRPX_DEVICE void st_stream(uint4* p, uint4 v) { __stcs(p, v); }

TRPX_DEVICE u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }

// ---- shared memory through 32-bit shared-window addresses ----
// The hot inner loops of the decoder address shared memory with ONE register per base (no generic-pointer
// arithmetic, no re-derivation of the window base): saddr() converts once, the accessors below take
// base + byte offset.  Loads are volatile asm (kept in program order with barriers) without a memory clobber.
typedef u32 saddr_t;
TRPX_DEVICE saddr_t saddr(const void* p) { return smem_addr(p); }
TRPX_DEVICE void* saddr_to_ptr(saddr_t a) { return __cvta_shared_to_generic((size_t)a); }
TRPX_DEVICE u32 lds_u32(saddr_t a)
{
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
template <int OFF>
TRPX_DEVICE u32 lds_u32_at(saddr_t a)                    // [a + OFF], OFF folded into the instruction
{
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
TRPX_DEVICE u32 lds_u16(saddr_t a)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
TRPX_DEVICE void sts_u32(saddr_t a, u32 x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
// (no memory clobber: for write-only staging that is read back only after a barrier)
TRPX_DEVICE void sts_u32_weak(saddr_t a, u32 x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x)); }
// the same, predicated inside the asm: straight-line code, no branch around the store
TRPX_DEVICE void sts_u32_if(bool c, saddr_t a, u32 x)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(a), "r"(x), "r"((u32)c));
}
TRPX_DEVICE void sts_v2(saddr_t a, u32 x, u32 y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
TRPX_DEVICE void sts_v4(saddr_t a, u32 x, u32 y, u32 z, u32 w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// the low n bits set, n clamped to 32 (one BMSK instead of shift / not / select)
TRPX_DEVICE u32 low_mask(u32 n)
{
    u32 m;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"(n));
    return m;
}

// ---- mbarrier (shared::cta) ----
TRPX_DEVICE void mbar_init(u64* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
TRPX_DEVICE void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
TRPX_DEVICE void mbar_arrive_expect_tx(u64* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
TRPX_DEVICE void mbar_arrive(u64* bar)    // release.cta: the arriving thread's earlier shared-memory writes are visible to waiters
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
TRPX_DEVICE void mbar_arrive_relaxed(u64* bar)   // no ordering of the thread's other memory operations
{
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
TRPX_DEVICE bool mbar_try_wait(u64* bar, u32 parity)
{
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
TRPX_DEVICE bool mbar_test(u64* bar, u32 parity) { return mbar_try_wait(bar, parity); }   // non-blocking: has that phase completed?
// 32 x 32 -> 64-bit product on the FMA pipe (IMAD.WIDE).  The bit packers shift by multiplying with a power of two:
// the ALU pipe (LOP3 / SHF / IADD3 / SEL / ISETP, half rate) is what bounds them, the FMA pipe is idle.  Inline PTX,
// because the compiler would turn a multiplication by (1 << n) back into shifts.
TRPX_DEVICE u64 mul_wide(u32 a, u32 b)
{
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
TRPX_DEVICE u32 mul_lo(u32 a, u32 b)           // a * b (IMAD), opaque to the strength reducer for the same reason
{
    u32 r;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
TRPX_DEVICE u32 mad_lo(u32 a, u32 b, u32 c)    // a * b + c (IMAD)
{
    u32 r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// Every wait in the kernels is bounded and NEVER traps (a trap poisons the CUDA context of the whole process): a
// wait that runs out raises an internal status word (>= ST_INTERNAL, the code names the wait) and gives up; every
// other wait sees that word and gives up too, so the kernel ends and the host reports TRPX_ERR_CUDA for the call.
constexpr u32 ST_INTERNAL = 0x100;
TRPX_DEVICE u64 now_ns()
{
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// limit_log2: the wait may last 2^limit_log2 microseconds (measured with %globaltimer every 256 polls)
struct WaitClock {
    u64 t0 = 0;
    TRPX_DEVICE bool expired(u32* status, u32 spins, u32 limit_log2, u32 code)
    {
        if ((spins & 255u) != 255u) return false;
        const u64 t = now_ns();
        if (!t0) t0 = t;
        if ((t - t0) >> (limit_log2 + 10)) {                 // the FIRST wait that runs out names the failure
            u32 old = *(volatile u32*)status;
            while (old < ST_INTERNAL) {
                const u32 prev = atomicCAS(status, old, ST_INTERNAL + code);
                if (prev == old) break;
                old = prev;
            }
        }
        return *(volatile u32*)status >= ST_INTERNAL;
    }
};
TRPX_DEVICE bool mbar_try_wait_hint(u64* bar, u32 parity, u32 hint_ns)
{
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
TRPX_DEVICE void mbar_wait(u64* bar, u32 parity, u32* status, u32 code)
{
    if (mbar_try_wait(bar, parity)) return;
    WaitClock wc;
    for (u32 spins = 0; !mbar_try_wait_hint(bar, parity, 1000u); ++spins) {   // (the hint alone does not stop the polling)
        if (wc.expired(status, spins, 21, code)) break;
        __nanosleep(100);
    }
}
TRPX_DEVICE void mbar_wait_sleep(u64* bar, u32 parity, u32* status, u32 code)
{
    WaitClock wc;
    for (u32 spins = 0; !mbar_try_wait_hint(bar, parity, 20000u); ++spins) {
        if (wc.expired(status, spins, 23, code)) break;
        __nanosleep(400);
    }
}

// ---- TMA bulk copies (1-D): SASS UBLKCP ----
// global -> shared, completion counted in bytes on `bar`; src/dst 16-byte aligned, bytes % 16 == 0
TRPX_DEVICE void bulk_g2s(void* smem_dst, const void* gsrc, u32 bytes, u64* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
// shared -> global (bulk async-group)
TRPX_DEVICE void bulk_s2g(void* gdst, const void* smem_src, u32 bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_addr(smem_src)), "r"(bytes)
                 : "memory");
}
TRPX_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk groups of this thread have finished READING shared memory
TRPX_DEVICE void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
TRPX_DEVICE void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (before bulk_s2g)
TRPX_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), u32 grid, u32 block, size_t smem, cudaStream_t st,
                          Args... args)
{
    if (smem > 48 * 1024) {   // opt in to large dynamic shared memory (up to 227 KB per CTA on sm_100a)
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, block, smem, st>>>(KArgs(args)...);
    return cudaGetLastError();
}

// grid-wide barrier of a cooperative launch (all CTAs resident by construction)
// (cooperative_groups aborts -- a trap -- when the grid was not launched cooperatively; launch_coop is the only way these
// kernels are launched, and checking validity here lets the compiler drop that path: no kernel of this library can trap)
TRPX_DEVICE void grid_sync()
{
    const cooperative_groups::grid_group g = cooperative_groups::this_grid();
    if (g.is_valid()) g.sync();
}
template <typename... P, typename... A>
inline cudaError_t launch_coop(void (*kern)(P...), u32 grid, u32 block, size_t smem, cudaStream_t st, A... args)
{
    std::tuple<P...> held{static_cast<P>(args)...};          // the launch reads the arguments through pointers
    void* kargs[sizeof...(P)];
    size_t n = 0;
    std::apply([&](auto&... a) { ((kargs[n++] = (void*)&a), ...); }, held);
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(block), kargs, smem, st);
}

#else
// ----------------------------------------------------------------------------- emulation backend
#define TRPX_DEVICE inline
#define TRPX_DEVICE_NOINLINE inline
#define TRPX_GRID_CONSTANT
#define TRPX_HD inline
#define TRPX_KERNEL
#define TRPX_SHARED static
#define TRPX_DYN_SMEM(name) unsigned char* name = ::emu::dyn_smem()
#define TRPX_LAUNCH_BOUNDS(t, b)

using ::emu::uint4;
using ::emu::make_uint4;
using ::emu::uint2;
using ::emu::make_uint2;

inline u32 tid() { return ::emu::cur().tid; }
inline u32 bid() { return ::emu::cur().bid; }
inline u32 nthreads() { return ::emu::cur().block_dim; }
inline u32 nblocks() { return ::emu::cur().grid_dim; }
inline void sync_block() { ::emu::sync_block(); }
inline void sync_warp() { ::emu::sync_warp(); }
inline void bar_sync(u32 id, u32 n) { ::emu::bar_sync(id, n); }
inline void bar_arrive(u32 id, u32 n) { ::emu::bar_arrive(id, n); }
inline void spin_hint() { ::emu::yield(); }
inline void trap() { ::emu::trap(); }

inline u32 shfl(u32 v, int src) { return (u32)::emu::shfl(v, src & 31); }
inline u64 shfl(u64 v, int src) { return ::emu::shfl(v, src & 31); }
inline u32 shfl_up(u32 v, int d) { return (u32)::emu::shfl_up(v, d); }
inline u64 shfl_up(u64 v, int d) { return ::emu::shfl_up(v, d); }
inline u32 shfl_down(u32 v, int d) { return (u32)::emu::shfl_down(v, d); }
inline u32 shfl_xor(u32 v, int m) { return (u32)::emu::shfl(v, (int)(::emu::cur().tid & 31) ^ m); }
inline u32 ballot(bool p) { return ::emu::ballot(p); }
inline bool all_lanes(bool p) { return ::emu::ballot(p) == ::emu::ballot(true); }
inline bool any_lane(bool p) { return ::emu::ballot(p) != 0; }
inline u32 warp_max(u32 v) { return ::emu::warp_reduce(v, 0); }
inline u32 warp_min_u32(u32 v) { return ~::emu::warp_reduce(~v, 0); }
inline u32 warp_or(u32 v) { return ::emu::warp_reduce(v, 1); }
inline u32 warp_add(u32 v) { return ::emu::warp_reduce(v, 2); }

inline int clz32(u32 x) { return x ? __builtin_clz(x) : 32; }
inline int clz64(u64 x) { return x ? __builtin_clzll(x) : 64; }
inline int ffs32(u32 x) { return __builtin_ffs((int)x); }
inline int ffs64(u64 x) { return __builtin_ffsll((long long)x); }
inline int popc32(u32 x) { return __builtin_popcount(x); }
inline u32 funnel_r(u32 lo, u32 hi, u32 sh)
{
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
inline u32 funnel_rc(u32 lo, u32 hi, u32 sh) { return sh >= 32 ? hi : (sh ? (lo >> sh) | (hi << (32 - sh)) : lo); }
inline u32 funnel_l(u32 lo, u32 hi, u32 sh)
{
    sh &= 31;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
inline u32 vabs2(u32 x)
{
    u32 r = 0;
    for (int k = 0; k < 2; ++k) {
        int16_t v = (int16_t)(x >> (16 * k));
        uint16_t a = (uint16_t)(v < 0 ? (uint16_t)(0u - (uint16_t)v) : (uint16_t)v);
        r |= (u32)a << (16 * k);
    }
    return r;
}
inline u32 vabs4(u32 x)
{
    u32 r = 0;
    for (int k = 0; k < 4; ++k) {
        int8_t v = (int8_t)(x >> (8 * k));
        uint8_t a = (uint8_t)(v < 0 ? (uint8_t)(0u - (uint8_t)v) : (uint8_t)v);
        r |= (u32)a << (8 * k);
    }
    return r;
}

inline u32 atomic_add(u32* p, u32 v) { u32 o = *p; *p = o + v; return o; }
inline u64 atomic_add(u64* p, u64 v) { u64 o = *p; *p = o + v; return o; }
inline u32 atomic_or(u32* p, u32 v) { u32 o = *p; *p = o | v; return o; }
inline u32 atomic_max(u32* p, u32 v) { u32 o = *p; if (v > o) *p = v; return o; }

inline u64 ld_relaxed(const u64* p) { return *(const volatile u64*)p; }
inline void st_relaxed(u64* p, u64 v) { *(volatile u64*)p = v; }
inline u32 ld_relaxed(const u32* p) { return *(const volatile u32*)p; }
// shared-window addresses: byte offsets into the block's dynamic shared memory (bounds-checked here)
typedef u32 saddr_t;
inline saddr_t saddr(const void* p) { return (u32)((const unsigned char*)p - ::emu::dyn_smem()); }
inline unsigned char* saddr_ptr(saddr_t a, u32 n)
{
    if ((size_t)a + n > 232448u || (a % n) != 0) ::emu::trap();
    return ::emu::dyn_smem() + a;
}
inline void* saddr_to_ptr(saddr_t a) { return ::emu::dyn_smem() + a; }
inline u32 lds_u32(saddr_t a) { return *(const u32*)saddr_ptr(a, 4); }
template <int OFF> inline u32 lds_u32_at(saddr_t a) { return *(const u32*)saddr_ptr(a + (u32)OFF, 4); }
inline u32 lds_u16(saddr_t a) { return *(const unsigned short*)saddr_ptr(a, 2); }
inline void sts_u32(saddr_t a, u32 x) { *(u32*)saddr_ptr(a, 4) = x; }
inline void sts_u32_weak(saddr_t a, u32 x) { *(u32*)saddr_ptr(a, 4) = x; }
inline void sts_u32_if(bool c, saddr_t a, u32 x) { if (c) *(u32*)saddr_ptr(a, 4) = x; }
inline void sts_v2(saddr_t a, u32 x, u32 y) { u32* d = (u32*)saddr_ptr(a, 8); d[0] = x; d[1] = y; }
inline void sts_v4(saddr_t a, u32 x, u32 y, u32 z, u32 w) { u32* d = (u32*)saddr_ptr(a, 16); d[0] = x; d[1] = y; d[2] = z; d[3] = w; }
inline u32 low_mask(u32 n) { return n >= 32 ? 0xffffffffu : (1u << n) - 1; }
inline u64 ldg_u64_opaque(const u64* p) { return *p; }
inline u32 ld_stream(const u32* p) { return *p; }
inline void st_stream(u32* p, u32 v) { *p = v; }
inline void st_stream(uint4* p, uint4 v) { *p = v; }

inline void mbar_init(u64* bar, u32 count) { ::emu::mbar_init(bar, count); }
inline void mbar_init_fence() {}
inline void mbar_arrive_expect_tx(u64* bar, u32 bytes) { ::emu::mbar_arrive_expect_tx(bar, bytes); }
inline void mbar_arrive(u64* bar) { ::emu::mbar_arrive(bar); }
inline void mbar_arrive_relaxed(u64* bar) { ::emu::mbar_arrive(bar); }
constexpr u32 ST_INTERNAL = 0x100;
struct WaitClock {
    bool expired(u32* status, u32 spins, u32 limit_log2, u32 code)
    {
        if (spins == (1u << 16)) fprintf(stderr, "emu: long wait: bid %u tid %u code %u (warp %u, round %u)\n", ::emu::cur().bid, ::emu::cur().tid, code & 15, (code >> 4) & 31, code >> 12);
        if (spins >> 18) { fprintf(stderr, "emu: wait %u (code %u, warp %u, round %u) expired\n", code, code & 15, (code >> 4) & 31, code >> 12); ::emu::trap(); }
        (void)limit_log2;
        (void)status;
        return false;
    }
};
inline void mbar_wait(u64* bar, u32 parity, u32*, u32 code) { ::emu::mbar_wait(bar, parity, code); }
inline bool mbar_test(u64* bar, u32 parity) { return ::emu::mbar_test(bar, parity); }
inline u64 mul_wide(u32 a, u32 b) { return (u64)a * b; }
inline u32 mul_lo(u32 a, u32 b) { return a * b; }
inline u32 mad_lo(u32 a, u32 b, u32 c) { return a * b + c; }
inline void mbar_wait_sleep(u64* bar, u32 parity, u32*, u32 code) { ::emu::mbar_wait(bar, parity, code); }
inline void bulk_g2s(void* d, const void* s, u32 bytes, u64* bar) { ::emu::bulk_g2s(d, s, bytes, bar); }
inline void bulk_s2g(void* d, const void* s, u32 bytes) { ::emu::bulk_s2g(d, s, bytes); }
inline void bulk_commit() {}
inline void bulk_wait_read0() {}
inline void bulk_wait_all0() {}
inline void fence_async_smem() {}

using ::emu::cudaError_t;
using ::emu::cudaStream_t;
using ::emu::cudaSuccess;

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), u32 grid, u32 block, size_t smem, cudaStream_t,
                          Args... args)
{
    ::emu::run_grid(grid, block, smem, [&]() { kern(KArgs(args)...); });
    return cudaSuccess;
}
// the emulator runs the blocks of a grid one after another, so a "cooperative" launch is one CTA
inline void grid_sync() { ::emu::sync_block(); }
template <typename... P, typename... A>
inline cudaError_t launch_coop(void (*kern)(P...), u32, u32 block, size_t smem, cudaStream_t, A... args)
{
    ::emu::run_grid(1, block, smem, [&]() { kern(static_cast<P>(args)...); });
    return cudaSuccess;
}
#endif

// ----------------------------------------------------------------------------- shared helpers
TRPX_HD u32 lane_of(u32 t) { return t & 31; }
TRPX_HD u32 warp_of(u32 t) { return t >> 5; }
TRPX_HD u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }

}  // namespace trpx
